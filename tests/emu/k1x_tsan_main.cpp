/* TEST INFRASTRUCTURE -- drives tests/emu/k1x_emu.cpp under -fsanitize=thread: a missing or misplaced barrier of the K1x
   kernel is a data race between the OS threads that stand for its CUDA threads (tests/test_k1x_emulation.py). */
#include <initializer_list>
#include <stdio.h>
#include <stdlib.h>
extern "C" int k1x_emu(int n, const float *windowed, float *gmdct, float *logfft, float *logmdct, float *ampmax);
int main() {
  for (int n : {2048, 256}) {
    static float in[2048], gm[1024], lf[1024], lm[1024]; float amp;
    for (int i = 0; i < n; i++) in[i] = (float)rand() / RAND_MAX - .5f;
    for (int rep = 0; rep < 3; rep++) k1x_emu(n, in, gm, lf, lm, &amp);
    printf("n=%d gm[5]=%g lf[7]=%g amp=%g\n", n, gm[5], lf[7], amp);
  }
  return 0;
}
