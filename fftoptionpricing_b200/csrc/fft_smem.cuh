// Batched fp64 FFT building blocks for sm_100a (replaces the serial radix-2 loop of fft.hpp:5-61).
//
// One transform of N = 2^n points (16 <= N <= 8192) lives in shared memory as double2 and is
// processed by N/16 threads; every thread owns 16 points per pass and does the butterflies in
// registers (radix 16 = 4 x 4; the last pass of a plan may be 8, 4 or 2 on 16 consecutive points).
//   plan(n): radices (16, 16, ..., r_last), strides S_1 = N/16, S_2 = S_1/16, ..., S_last = 1.
// The passes are IN PLACE.  Decimation in frequency (dif_*) maps natural-order input to
// digit-reversed output:  X[k] ends at position  pos(k) = t_1 S_1 + t_2 S_2 + ... + t_p,
// k = t_1 + r_1 (t_2 + r_2 (t_3 + ...)).  Decimation in time (dit_*) is the exact transpose: it
// takes digit-reversed input to natural-order output.  FFT -> pointwise multiply -> IFFT therefore
// needs no reordering at all (forward DIF, multiply in digit-reversed order, inverse DIT), and the
// last forward pass / multiply / first inverse pass touch the same 16 registers of each thread.
//
// Shared-memory index padding  PAD(i) = i + (i >> 4)  makes every pass of the 16^k plans
// conflict-free for 16-byte accesses (stride-1 pass: lane stride 17 double2 = 272 B).
//
// Twiddles are exact: W_N^k = e^{-2 pi i k/N} from a table filled with sincospi (never by
// recurrence -- SURVEY.md section 0 shows the reference's own recurrence costs it 1e-13..1e-12).
// Callers either read the table (__ldg, L1/L2 resident) or, in the FSTS time loop, keep the base
// twiddle of each pass in registers and form its 15 powers by a depth-4 multiplication tree.
//
// All functions are __host__ __device__ so the index algebra is unit-tested on the CPU
// (tests/test_fft_host.py builds csrc/host_fft_test.cu); there is no CPU product path.
#pragma once
#include "cplx.cuh"

#ifdef __CUDA_ARCH__
#define FOP_LDG(p) __ldg(p)
#else
#define FOP_LDG(p) (*(p))
#endif

namespace fop {

#define FOP_HD __host__ __device__ __forceinline__

FOP_HD int fft_pad(int i) { return i + (i >> 4); }
FOP_HD int fft_padded_size(int N) { return N + (N >> 4) + 1; }

// number of passes and radix of the last pass for N = 2^n, n >= 4
FOP_HD int fft_last_radix(int n) { return (n & 3) == 0 ? 16 : (1 << (n & 3)); }
FOP_HD int fft_num_passes(int n) { return (n + 3) >> 2; }

// position of X[k] after the DIF passes (mixed-radix digit reversal); an involution only for
// uniform radices, so fft_inv_pos is provided separately.
FOP_HD int fft_pos(int k, int n) {
  int S = 1 << n, pos = 0;
  int rem = n;
  while (rem > 0) {
    const int bits = rem >= 4 ? 4 : rem;  // radix 16 while possible, then the last radix
    // plans put the small radix LAST, i.e. radices are 16,16,...,r_last
    const int rb = (rem > 4 || (rem & 3) == 0) ? 4 : bits;
    S >>= rb;
    pos += (k & ((1 << rb) - 1)) * S;
    k >>= rb;
    rem -= rb;
  }
  return pos;
}
// inverse map: which k lives at position `pos`
FOP_HD int fft_inv_pos(int pos, int n) {
  int S = 1 << n, k = 0, shift = 0;
  int rem = n;
  while (rem > 0) {
    const int rb = (rem > 4 || (rem & 3) == 0) ? 4 : rem;
    S >>= rb;
    const int t = (pos / S) & ((1 << rb) - 1);
    k |= t << shift;
    shift += rb;
    rem -= rb;
  }
  return k;
}

// ---------------------------------------------------------------- register butterflies
// multiply by SIGN * i
template <int SIGN>
FOP_HD cd mul_si(cd a) { return SIGN > 0 ? mk(-a.y, a.x) : mk(a.y, -a.x); }

// 4-point DFT, natural order in and out, w = e^{SIGN i pi/2}
template <int SIGN>
FOP_HD void dft4(cd& a0, cd& a1, cd& a2, cd& a3) {
  const cd t0 = a0 + a2, t1 = a0 - a2, t2 = a1 + a3, t3 = mul_si<SIGN>(a1 - a3);
  a0 = t0 + t2;
  a2 = t0 - t2;
  a1 = t1 + t3;
  a3 = t1 - t3;
}
FOP_HD void dft2(cd& a0, cd& a1) {
  const cd t = a0 - a1;
  a0 = a0 + a1;
  a1 = t;
}
// multiply by e^{SIGN i pi m/8}, m a compile-time constant in [0,16)
template <int SIGN, int M>
FOP_HD cd mul_w16(cd a) {
  constexpr double C1 = 0.92387953251128675613;  // cos(pi/8)
  constexpr double S1 = 0.38268343236508977173;  // sin(pi/8)
  constexpr double H = 0.70710678118654752440;   // sqrt(1/2)
  constexpr double s = SIGN > 0 ? 1.0 : -1.0;
  if (M == 0) return a;
  if (M == 1) return mk(C1 * a.x - s * S1 * a.y, C1 * a.y + s * S1 * a.x);
  if (M == 2) return mk(H * (a.x - s * a.y), H * (a.y + s * a.x));
  if (M == 3) return mk(S1 * a.x - s * C1 * a.y, S1 * a.y + s * C1 * a.x);
  if (M == 4) return mul_si<SIGN>(a);
  if (M == 6) return mk(-H * (a.x + s * a.y), H * (s * a.x - a.y));
  if (M == 9) return mk(-C1 * a.x + s * S1 * a.y, -C1 * a.y - s * S1 * a.x);
  return a;
}

// In-register DFT of length R.  On return the k-th output sits in x[Dft<R>::out(k)].
template <int R, int SIGN>
struct Dft;

template <int SIGN>
struct Dft<2, SIGN> {
  static FOP_HD void run(cd* x) { dft2(x[0], x[1]); }
  static FOP_HD constexpr int out(int k) { return k; }
};
template <int SIGN>
struct Dft<4, SIGN> {
  static FOP_HD void run(cd* x) { dft4<SIGN>(x[0], x[1], x[2], x[3]); }
  static FOP_HD constexpr int out(int k) { return k; }
};
// 8 = 4 (over n1) x 2 (over n0), n = 2 n1 + n0, k = k0 + 4 k1
template <int SIGN>
struct Dft<8, SIGN> {
  static FOP_HD void run(cd* x) {
    dft4<SIGN>(x[0], x[2], x[4], x[6]);  // n0 = 0: a[0][k0] in x[2 k0]
    dft4<SIGN>(x[1], x[3], x[5], x[7]);  // n0 = 1: a[1][k0] in x[2 k0 + 1]
    x[3] = mul_w16<SIGN, 2>(x[3]);       // W8^{k0}, k0 = 1
    x[5] = mul_w16<SIGN, 4>(x[5]);       // k0 = 2
    x[7] = mul_w16<SIGN, 6>(x[7]);       // k0 = 3
    dft2(x[0], x[1]);                    // y[k0 + 4 k1] in x[2 k0 + k1]
    dft2(x[2], x[3]);
    dft2(x[4], x[5]);
    dft2(x[6], x[7]);
  }
  static FOP_HD constexpr int out(int k) { return 2 * (k & 3) + (k >> 2); }
};
// 16 = 4 x 4, n = 4 n1 + n0, k = k0 + 4 k1
template <int SIGN>
struct Dft<16, SIGN> {
  static FOP_HD void run(cd* x) {
    dft4<SIGN>(x[0], x[4], x[8], x[12]);   // a[n0][k0] lands in x[4 k0 + n0]
    dft4<SIGN>(x[1], x[5], x[9], x[13]);
    dft4<SIGN>(x[2], x[6], x[10], x[14]);
    dft4<SIGN>(x[3], x[7], x[11], x[15]);
    x[5] = mul_w16<SIGN, 1>(x[5]);    // (n0,k0) = (1,1)
    x[6] = mul_w16<SIGN, 2>(x[6]);    // (2,1)
    x[7] = mul_w16<SIGN, 3>(x[7]);    // (3,1)
    x[9] = mul_w16<SIGN, 2>(x[9]);    // (1,2)
    x[10] = mul_w16<SIGN, 4>(x[10]);  // (2,2)
    x[11] = mul_w16<SIGN, 6>(x[11]);  // (3,2)
    x[13] = mul_w16<SIGN, 3>(x[13]);  // (1,3)
    x[14] = mul_w16<SIGN, 6>(x[14]);  // (2,3)
    x[15] = mul_w16<SIGN, 9>(x[15]);  // (3,3)
    dft4<SIGN>(x[0], x[1], x[2], x[3]);  // y[k0 + 4 k1] in x[4 k0 + k1]
    dft4<SIGN>(x[4], x[5], x[6], x[7]);
    dft4<SIGN>(x[8], x[9], x[10], x[11]);
    dft4<SIGN>(x[12], x[13], x[14], x[15]);
  }
  static FOP_HD constexpr int out(int k) { return 4 * (k & 3) + (k >> 2); }
};

// twiddle e^{SIGN 2 pi i idx / N} from the forward table tw[idx] = e^{-2 pi i idx/N}
template <int SIGN>
FOP_HD cd tw_load(const double2* __restrict__ tw, int idx) {
  const double2 w = FOP_LDG(tw + idx);
  return SIGN < 0 ? mk(w.x, w.y) : mk(w.x, -w.y);
}

// w^0 .. w^15 by a depth-4 multiplication tree (p[0] unused = 1)
FOP_HD void tw_powers16(cd w, cd* p) {
  p[1] = w;
  p[2] = w * w;
  p[3] = p[2] * w;
  p[4] = p[2] * p[2];
  p[5] = p[4] * w;
  p[6] = p[3] * p[3];
  p[7] = p[4] * p[3];
  p[8] = p[4] * p[4];
  p[9] = p[8] * w;
  p[10] = p[5] * p[5];
  p[11] = p[8] * p[3];
  p[12] = p[6] * p[6];
  p[13] = p[8] * p[5];
  p[14] = p[7] * p[7];
  p[15] = p[8] * p[7];
}

// ---------------------------------------------------------------- thread-level pass pieces
// The 16 registers x[0..15] of thread b hold, for a radix-16 pass of stride S:
//   positions base + t S,  base = (b / S) * 16 S + (b % S),  q = b % S.
FOP_HD int fft_base16(int b, int S) { return (b / S) * (S << 4) + (b % S); }

// radix-16 butterfly + DIF output twiddles  y_k *= W_L^{q k}, L = 16 S  (table stride N/L)
template <int SIGN>
FOP_HD void dif16_regs(cd* x, int q, int twstride, const double2* __restrict__ tw) {
  Dft<16, SIGN>::run(x);
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    const int r = Dft<16, SIGN>::out(k);
    x[r] = x[r] * tw_load<SIGN>(tw, q * k * twstride);
  }
}
// DIT: input twiddles then butterfly
template <int SIGN>
FOP_HD void dit16_regs(cd* x, int q, int twstride, const double2* __restrict__ tw) {
#pragma unroll
  for (int t = 1; t < 16; ++t) x[t] = x[t] * tw_load<SIGN>(tw, q * t * twstride);
  Dft<16, SIGN>::run(x);
}
// same with the base twiddle w = W_L^{q} (sign already applied) in a register
template <int SIGN>
FOP_HD void dif16_regs_pow(cd* x, cd w) {
  Dft<16, SIGN>::run(x);
  cd p[16];
  tw_powers16(w, p);
#pragma unroll
  for (int k = 1; k < 16; ++k) {
    const int r = Dft<16, SIGN>::out(k);
    x[r] = x[r] * p[k];
  }
}
template <int SIGN>
FOP_HD void dit16_regs_pow(cd* x, cd w) {
  cd p[16];
  tw_powers16(w, p);
#pragma unroll
  for (int t = 1; t < 16; ++t) x[t] = x[t] * p[t];
  Dft<16, SIGN>::run(x);
}

// last pass of a plan on the 16 CONSECUTIVE points of a thread (stride 1, no twiddles):
// 16/RL butterflies of radix RL; x[i] <-> position 16 b + i.
template <int RL, int SIGN>
FOP_HD void last_pass_regs(cd* x) {
#pragma unroll
  for (int i = 0; i < 16 / RL; ++i) {
    cd y[RL];
#pragma unroll
    for (int t = 0; t < RL; ++t) y[t] = x[i * RL + t];
    Dft<RL, SIGN>::run(y);
#pragma unroll
    for (int k = 0; k < RL; ++k) x[i * RL + k] = y[Dft<RL, SIGN>::out(k)];
  }
}

// shared-memory <-> register moves for a radix-16 pass (stride S) and for the consecutive pass
// When S is a multiple of 16, pad(base + t S) = pad(base) + t (S + S/16): one padded base plus
// constant offsets (immediates once S is a compile-time constant).
FOP_HD int fft_pad_off(int base, int t, int S) {
  return (S & 15) == 0 ? fft_pad(base) + t * (S + (S >> 4)) : fft_pad(base + t * S);
}
FOP_HD void load16(const cd* sm, int base, int S, cd* x) {
#pragma unroll
  for (int t = 0; t < 16; ++t) x[t] = sm[fft_pad_off(base, t, S)];
}
// stores the k-th output (held in x[perm(k)]) to base + k S
template <int SIGN>
FOP_HD void store16_dft(cd* sm, int base, int S, const cd* x) {
#pragma unroll
  for (int k = 0; k < 16; ++k) sm[fft_pad_off(base, k, S)] = x[Dft<16, SIGN>::out(k)];
}
FOP_HD void store16(cd* sm, int base, int S, const cd* x) {
#pragma unroll
  for (int t = 0; t < 16; ++t) sm[fft_pad_off(base, t, S)] = x[t];
}
// after dit16 the outputs are also permuted: natural k-th output in x[out(k)]
template <int SIGN>
FOP_HD void unpermute16(cd* x) {
  cd y[16];
#pragma unroll
  for (int k = 0; k < 16; ++k) y[k] = x[Dft<16, SIGN>::out(k)];
#pragma unroll
  for (int k = 0; k < 16; ++k) x[k] = y[k];
}

// ---------------------------------------------------------------- whole transforms
// One pass for one thread; b in [0, N/16), N = 2^n >= 16, p in [0, fft_num_passes(n)).
// DIF: natural-order input -> digit-reversed output (passes run p = 0, 1, ...).
template <int SIGN>
FOP_HD void fft_dif_pass(cd* sm, int b, int n, int p, const double2* __restrict__ tw) {
  const int N = 1 << n;
  const int rl = fft_last_radix(n);
  const int n16 = fft_num_passes(n) - (rl == 16 ? 0 : 1);  // number of radix-16 passes
  cd x[16];
  if (p < n16) {
    const int S = N >> (4 * (p + 1));
    const int base = fft_base16(b, S), q = b % S;
    load16(sm, base, S, x);
    // one exact table twiddle W_L^q per thread, its 15 powers by multiplication (depth 4)
    if (S > 1) dif16_regs_pow<SIGN>(x, tw_load<SIGN>(tw, q * (N / (S << 4))));
    else Dft<16, SIGN>::run(x);
    store16_dft<SIGN>(sm, base, S, x);
  } else {
    load16(sm, 16 * b, 1, x);
    if (rl == 8) last_pass_regs<8, SIGN>(x);
    else if (rl == 4) last_pass_regs<4, SIGN>(x);
    else last_pass_regs<2, SIGN>(x);
    store16(sm, 16 * b, 1, x);
  }
}
// DIT: digit-reversed input -> natural-order output; the exact transpose of the DIF passes, run
// in the opposite order (small last radix first, then strides rl, 16 rl, ...).
template <int SIGN>
FOP_HD void fft_dit_pass(cd* sm, int b, int n, int p, const double2* __restrict__ tw) {
  const int N = 1 << n;
  const int rl = fft_last_radix(n);
  cd x[16];
  if (rl != 16 && p == 0) {
    load16(sm, 16 * b, 1, x);
    if (rl == 8) last_pass_regs<8, SIGN>(x);
    else if (rl == 4) last_pass_regs<4, SIGN>(x);
    else last_pass_regs<2, SIGN>(x);
    store16(sm, 16 * b, 1, x);
    return;
  }
  const int p16 = rl == 16 ? p : p - 1;  // index among the radix-16 passes
  const int S = (rl == 16 ? 1 : rl) << (4 * p16);
  const int base = fft_base16(b, S), q = b % S;
  load16(sm, base, S, x);
  if (S > 1) dit16_regs_pow<SIGN>(x, tw_load<SIGN>(tw, q * (N / (S << 4))));
  else Dft<16, SIGN>::run(x);
  store16_dft<SIGN>(sm, base, S, x);
}

// `sync` is a functor (e.g. [] { __syncthreads(); }); every thread of the group calls it the same
// number of times (once after every pass).
template <int SIGN, typename Sync>
FOP_HD void fft_dif_smem(cd* sm, int b, int n, const double2* __restrict__ tw, Sync sync) {
  const int np = fft_num_passes(n);
  for (int p = 0; p < np; ++p) {
    fft_dif_pass<SIGN>(sm, b, n, p, tw);
    sync();
  }
}
template <int SIGN, typename Sync>
FOP_HD void fft_dit_smem(cd* sm, int b, int n, const double2* __restrict__ tw, Sync sync) {
  const int np = fft_num_passes(n);
  for (int p = 0; p < np; ++p) {
    fft_dit_pass<SIGN>(sm, b, n, p, tw);
    sync();
  }
}

}  // namespace fop
