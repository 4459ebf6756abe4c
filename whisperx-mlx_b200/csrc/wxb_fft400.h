// wxb_fft400.h — 400-point complex Stockham FFT (radix 4,4,5,5) written so the same code runs
// inside the log-mel CUDA kernel and in a host emulation (tests/host_emu/fft400_emu.cpp), which
// walks the "threads" sequentially between barriers.  Two real frames ride in one complex
// transform (frame a = real part, frame b = imaginary part) and are separated afterwards.
//
// Pass p works on sub-transform size Ns (product of the radices already applied).  Butterfly i
// (0 <= i < 400/R) of one transform reads in[i + q*400/R], q < R, multiplies by the twiddle
// w^(q*k*400/(Ns*R)) with k = i % Ns, does an R-point DFT and writes out[(i-k)*R + k + q*Ns].
#pragma once
#ifndef WXB_HD
#ifdef __CUDACC__
#define WXB_HD __host__ __device__ __forceinline__
#else
#define WXB_HD inline
#endif
#endif

struct cpx {
  float x, y;
};
WXB_HD cpx cadd(cpx a, cpx b) { return cpx{a.x + b.x, a.y + b.y}; }
WXB_HD cpx csub(cpx a, cpx b) { return cpx{a.x - b.x, a.y - b.y}; }
WXB_HD cpx cmul(cpx a, cpx b) { return cpx{a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
WXB_HD cpx cscale(cpx a, float s) { return cpx{a.x * s, a.y * s}; }
WXB_HD cpx mul_neg_i(cpx a) { return cpx{a.y, -a.x}; }  // a * (-i)
WXB_HD cpx mul_pos_i(cpx a) { return cpx{-a.y, a.x}; }  // a * (+i)

#define WXB_FFT_N 400

WXB_HD void dft4(cpx* v) {
  cpx a = cadd(v[0], v[2]), b = csub(v[0], v[2]);
  cpx c = cadd(v[1], v[3]), d = mul_neg_i(csub(v[1], v[3]));
  v[0] = cadd(a, c);
  v[1] = cadd(b, d);
  v[2] = csub(a, c);
  v[3] = csub(b, d);
}

WXB_HD void dft5(cpx* v) {
  const float c1 = 0.30901699437494742f;   // cos(2pi/5)
  const float c2 = -0.80901699437494742f;  // cos(4pi/5)
  const float s1 = 0.95105651629515357f;   // sin(2pi/5)
  const float s2 = 0.58778525229247313f;   // sin(4pi/5)
  cpx a1 = cadd(v[1], v[4]), a2 = cadd(v[2], v[3]);
  cpx b1 = csub(v[1], v[4]), b2 = csub(v[2], v[3]);
  cpx t1 = cadd(v[0], cadd(cscale(a1, c1), cscale(a2, c2)));
  cpx t2 = cadd(v[0], cadd(cscale(a1, c2), cscale(a2, c1)));
  cpx u1 = cadd(cscale(b1, s1), cscale(b2, s2));
  cpx u2 = csub(cscale(b1, s2), cscale(b2, s1));
  v[0] = cadd(v[0], cadd(a1, a2));
  v[1] = cadd(t1, mul_neg_i(u1));
  v[4] = cadd(t1, mul_pos_i(u1));
  v[2] = cadd(t2, mul_neg_i(u2));
  v[3] = cadd(t2, mul_pos_i(u2));
}

// One butterfly of one pass.  `tw` is the table tw[k] = exp(-2*pi*i*k/400), k < 400.
template <int R, int NS>
WXB_HD void fft400_butterfly(const cpx* in, cpx* out, const cpx* tw, int i) {
  constexpr int T = WXB_FFT_N / R;
  constexpr int TWSTEP = WXB_FFT_N / (NS * R);
  const int k = i % NS;
  cpx v[R];
#pragma unroll
  for (int q = 0; q < R; ++q) v[q] = in[i + q * T];
  if (NS > 1) {
#pragma unroll
    for (int q = 1; q < R; ++q) v[q] = cmul(v[q], tw[q * k * TWSTEP]);
  }
  if (R == 4) dft4(v); else dft5(v);
  const int j = (i - k) * R + k;
#pragma unroll
  for (int q = 0; q < R; ++q) out[j + q * NS] = v[q];
}
