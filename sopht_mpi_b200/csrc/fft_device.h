// In-kernel FFT building blocks for the pruned Poisson pipeline.
//
// A line of length n = 2^k is transformed by T = n/16 threads; every thread keeps 16
// complex points in registers at positions  t + p*T  (p = 0..15).  That position set is the
// same for every Stockham stage, so global loads, global stores and the shared-memory
// re-reads are all unit-stride across threads, and zero-padding pruning is "skip p >= 8".
// Stages of radix 16/8/4/2 run entirely in registers; between stages the line is exchanged
// through (padded) shared memory.  Index algebra prototyped in tools/fft_prototype.py.
#pragma once
#include "sb200_rt.h"

template <typename T>
struct alignas(2 * sizeof(T)) C2 {
  T x, y;
};

template <typename T>
SB_HD C2<T> cmul(C2<T> a, C2<T> b) {
  return C2<T>{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x};
}
template <typename T>
SB_HD C2<T> cadd(C2<T> a, C2<T> b) { return C2<T>{a.x + b.x, a.y + b.y}; }
template <typename T>
SB_HD C2<T> csub(C2<T> a, C2<T> b) { return C2<T>{a.x - b.x, a.y - b.y}; }
template <typename T>
SB_HD C2<T> cconj(C2<T> a) { return C2<T>{a.x, -a.y}; }
// multiply by -i  (forward quarter turn)
template <typename T>
SB_HD C2<T> cmul_mi(C2<T> a) { return C2<T>{a.y, -a.x}; }

#define SB_FFT_R 16        // points per thread
#define SB_FFT_MAXSTAGES 4

struct SbFftPlan {
  int n;        // transform length (power of two, >= 16)
  int threads;  // n / 16 threads per line
  int nstages;
  int radix[SB_FFT_MAXSTAGES];
};

static inline int sb_fft_make_plan(int n, SbFftPlan* p) {
  if (n < 16 || (n & (n - 1))) return -1;
  p->n = n;
  p->threads = n / SB_FFT_R;
  p->nstages = 0;
  int m = n;
  while (m > 1) {
    int r = m >= 16 ? 16 : m;
    if (p->nstages >= SB_FFT_MAXSTAGES) return -1;
    p->radix[p->nstages++] = r;
    m /= r;
  }
  return 0;
}

// ---------------------------------------------------------------- small DFTs (forward)
template <typename T>
SB_D void dft2(C2<T>& a, C2<T>& b) {
  const C2<T> s = cadd(a, b), d = csub(a, b);
  a = s;
  b = d;
}
template <typename T>
SB_D void dft4(C2<T>& a0, C2<T>& a1, C2<T>& a2, C2<T>& a3) {
  const C2<T> t0 = cadd(a0, a2), t1 = csub(a0, a2), t2 = cadd(a1, a3), t3 = cmul_mi(csub(a1, a3));
  a0 = cadd(t0, t2);
  a1 = cadd(t1, t3);
  a2 = csub(t0, t2);
  a3 = csub(t1, t3);
}
// a[0..R) -> DFT in natural order, R in {2,4,8,16}
template <typename T, int R>
SB_D void dft_small(C2<T>* a) {
  if (R == 2) {
    dft2(a[0], a[1]);
  } else if (R == 4) {
    dft4(a[0], a[1], a[2], a[3]);
  } else if (R == 8) {
    // 8 = 4 (inner, over n1, stride 2) x 2 (outer, over n2)
    const T h = T(0.70710678118654752440);
    C2<T> e[4] = {a[0], a[2], a[4], a[6]}, o[4] = {a[1], a[3], a[5], a[7]};
    dft4(e[0], e[1], e[2], e[3]);
    dft4(o[0], o[1], o[2], o[3]);
    // twiddle W8^k1 on the odd branch
    o[1] = C2<T>{h * (o[1].x + o[1].y), h * (o[1].y - o[1].x)};
    o[2] = cmul_mi(o[2]);
    o[3] = C2<T>{h * (o[3].y - o[3].x), -h * (o[3].x + o[3].y)};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      a[k] = cadd(e[k], o[k]);
      a[k + 4] = csub(e[k], o[k]);
    }
  } else {
    // 16 = 4 (inner over n1, stride 4) x 4 (outer over n2); X[k1 + 4 k2]
    const T c1 = T(0.92387953251128675613), s1 = T(0.38268343236508977173);
    const T h = T(0.70710678118654752440);
    C2<T> b[4][4];  // b[n2][k1]
#pragma unroll
    for (int n2 = 0; n2 < 4; ++n2) {
      b[n2][0] = a[n2];
      b[n2][1] = a[n2 + 4];
      b[n2][2] = a[n2 + 8];
      b[n2][3] = a[n2 + 12];
      dft4(b[n2][0], b[n2][1], b[n2][2], b[n2][3]);
    }
    // twiddles W16^(n2*k1) = exp(-2 pi i n2 k1 / 16)
    const C2<T> w1{c1, -s1}, w2{h, -h}, w3{s1, -c1}, w6{-h, -h}, w9{-c1, s1};
    b[1][1] = cmul(b[1][1], w1);
    b[1][2] = cmul(b[1][2], w2);
    b[1][3] = cmul(b[1][3], w3);
    b[2][1] = cmul(b[2][1], w2);
    b[2][2] = cmul_mi(b[2][2]);  // W16^4 = -i
    b[2][3] = cmul(b[2][3], w6);
    b[3][1] = cmul(b[3][1], w3);
    b[3][2] = cmul(b[3][2], w6);
    b[3][3] = cmul(b[3][3], w9);
#pragma unroll
    for (int k1 = 0; k1 < 4; ++k1) {
      dft4(b[0][k1], b[1][k1], b[2][k1], b[3][k1]);
      a[k1] = b[0][k1];
      a[k1 + 4] = b[1][k1];
      a[k1 + 8] = b[2][k1];
      a[k1 + 12] = b[3][k1];
    }
  }
}

// padded shared-memory index of element i of local line l
template <bool LINE_FASTEST>
SB_D int sb_sidx(int l, int i, int lines, int npad) {
  const int ip = i + (i >> 4);
  return LINE_FASTEST ? ip * lines + l : l * npad + ip;
}
SB_HD int sb_fft_npad(int n) { return n + (n >> 4) + 1; }

// One radix-R stage for the butterflies owned by thread t.
//   v[p] <-> element t + p*T.  Ns = product of the radices already applied.
// last == false: results go to shared memory (natural Stockham positions), caller syncs and
// re-reads; last == true: results return to v[] (their natural positions coincide with p).
template <typename T, int R, bool LINE_FASTEST>
SB_D void sb_fft_stage(C2<T>* v, int t, int threads, int n, int Ns, bool last, const C2<T>* __restrict__ tw,
                       C2<T>* sm, int l, int lines, int npad) {
  constexpr int M = SB_FFT_R / R;  // butterflies per thread
#pragma unroll
  for (int m = 0; m < M; ++m) {
    C2<T> a[R];
#pragma unroll
    for (int q = 0; q < R; ++q) a[q] = v[m + M * q];
    const int j = t + m * threads;
    const int k = j % Ns;  // Ns is a power of two
    if (Ns > 1) {
      const int step = k * (n / (Ns * R));
#pragma unroll
      for (int q = 1; q < R; ++q) a[q] = cmul(a[q], tw[step * q]);
    }
    dft_small<T, R>(a);
    if (last) {
#pragma unroll
      for (int q = 0; q < R; ++q) v[m + M * q] = a[q];
    } else {
      const int j0 = (j / Ns) * Ns * R + k;
#pragma unroll
      for (int q = 0; q < R; ++q) sm[sb_sidx<LINE_FASTEST>(l, j0 + q * Ns, lines, npad)] = a[q];
    }
  }
}

// Forward FFT of the line held in v[] (all threads of the block must call this; the
// __syncthreads inside are block-wide).  On return v[p] = X[t + p*T].
template <typename T, bool LINE_FASTEST>
SB_D void sb_fft_forward(C2<T>* v, const SbFftPlan& plan, int t, const C2<T>* __restrict__ tw, C2<T>* sm,
                         int l, int lines, int npad) {
  int Ns = 1;
  for (int s = 0; s < plan.nstages; ++s) {
    const int r = plan.radix[s];
    const bool last = s == plan.nstages - 1;
    if (r == 16)
      sb_fft_stage<T, 16, LINE_FASTEST>(v, t, plan.threads, plan.n, Ns, last, tw, sm, l, lines, npad);
    else if (r == 8)
      sb_fft_stage<T, 8, LINE_FASTEST>(v, t, plan.threads, plan.n, Ns, last, tw, sm, l, lines, npad);
    else if (r == 4)
      sb_fft_stage<T, 4, LINE_FASTEST>(v, t, plan.threads, plan.n, Ns, last, tw, sm, l, lines, npad);
    else
      sb_fft_stage<T, 2, LINE_FASTEST>(v, t, plan.threads, plan.n, Ns, last, tw, sm, l, lines, npad);
    if (!last) {
      __syncthreads();
#pragma unroll
      for (int p = 0; p < SB_FFT_R; ++p)
        v[p] = sm[sb_sidx<LINE_FASTEST>(l, t + p * plan.threads, lines, npad)];
      __syncthreads();
    }
    Ns *= r;
  }
}

// Unnormalised inverse through conj(fft(conj(x))).
template <typename T, bool LINE_FASTEST>
SB_D void sb_fft_inverse(C2<T>* v, const SbFftPlan& plan, int t, const C2<T>* __restrict__ tw, C2<T>* sm,
                         int l, int lines, int npad) {
#pragma unroll
  for (int p = 0; p < SB_FFT_R; ++p) v[p].y = -v[p].y;
  sb_fft_forward<T, LINE_FASTEST>(v, plan, t, tw, sm, l, lines, npad);
#pragma unroll
  for (int p = 0; p < SB_FFT_R; ++p) v[p].y = -v[p].y;
}
