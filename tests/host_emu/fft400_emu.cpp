// Host emulation of the log-mel kernel's FFT core (whisperx-mlx_b200/csrc/wxb_fft400.h):
// runs the four Stockham passes butterfly by butterfly and checks against a naive fp64 DFT,
// then checks the two-real-frames-in-one-complex-FFT separation.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "../../whisperx-mlx_b200/csrc/wxb_fft400.h"

int main() {
  const int N = WXB_FFT_N;
  std::vector<cpx> tw(N), a(N), b(N);
  for (int k = 0; k < N; ++k) {
    double ang = -2.0 * M_PI * k / N;
    tw[k] = cpx{(float)cos(ang), (float)sin(ang)};
  }
  srand(7);
  std::vector<double> fa(N), fb(N);
  for (int n = 0; n < N; ++n) {
    fa[n] = rand() / (double)RAND_MAX - 0.5;
    fb[n] = rand() / (double)RAND_MAX - 0.5;
    a[n] = cpx{(float)fa[n], (float)fb[n]};
  }
  for (int i = 0; i < N / 4; ++i) fft400_butterfly<4, 1>(a.data(), b.data(), tw.data(), i);
  for (int i = 0; i < N / 4; ++i) fft400_butterfly<4, 4>(b.data(), a.data(), tw.data(), i);
  for (int i = 0; i < N / 5; ++i) fft400_butterfly<5, 16>(a.data(), b.data(), tw.data(), i);
  for (int i = 0; i < N / 5; ++i) fft400_butterfly<5, 80>(b.data(), a.data(), tw.data(), i);
  double maxerr = 0, maxmag = 0, maxerr_sep = 0;
  for (int f = 0; f <= 200; ++f) {
    double zr = 0, zi = 0, ar = 0, ai = 0, br = 0, bi = 0;
    for (int n = 0; n < N; ++n) {
      double ang = -2.0 * M_PI * f * n / N, c = cos(ang), s = sin(ang);
      zr += fa[n] * c - fb[n] * s;
      zi += fa[n] * s + fb[n] * c;
      ar += fa[n] * c; ai += fa[n] * s;
      br += fb[n] * c; bi += fb[n] * s;
    }
    maxerr = fmax(maxerr, fmax(fabs(zr - a[f].x), fabs(zi - a[f].y)));
    maxmag = fmax(maxmag, hypot(zr, zi));
    // separation: Xa = (Z[f] + conj(Z[N-f]))/2 ; Xb = (Z[f] - conj(Z[N-f]))/(2i)
    cpx z = a[f], zc = a[(N - f) % N];
    double xar = 0.5 * (z.x + zc.x), xai = 0.5 * (z.y - zc.y);
    double dr = z.x - zc.x, di = z.y + zc.y;  // Z - conj(Zc)
    double xbr = 0.5 * di, xbi = -0.5 * dr;   // divide by 2i
    maxerr_sep = fmax(maxerr_sep, fmax(fmax(fabs(xar - ar), fabs(xai - ai)), fmax(fabs(xbr - br), fabs(xbi - bi))));
  }
  printf("fft400 max abs err %.3e (max |Z| %.3f), separation err %.3e\n", maxerr, maxmag, maxerr_sep);
  if (maxerr > 2e-5 || maxerr_sep > 2e-5) { printf("FAIL\n"); return 1; }
  printf("OK\n");
  return 0;
}
