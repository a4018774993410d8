/* Checker for csrc/env_core.cuh::div_by_const: the FMA-refined x * RN(1/h) against the IEEE division x / h for EVERY fp32
 * x with 2^-31 <= |x| < 2^24 (both signs).  Prints the number of mismatches per divisor.  usage: div_const_check h1 h2 ... */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
static inline float divc(float x, float h, float y) {
  float q = x * y;
  float r = fmaf(-q, h, x);
  q = fmaf(r, y, q);
  r = fmaf(-q, h, x);
  return fmaf(r, y, q);
}
int main(int argc, char** argv) {
  for (int k = 1; k < argc; ++k) {
    const float h = (float)atof(argv[k]), y = 1.0f / h;
    long bad = 0;
    for (uint32_t bits = 0x30000000u; bits < 0x4B800000u; ++bits) {
      float x;
      memcpy(&x, &bits, 4);
      if (x / h != divc(x, h, y)) ++bad;
      if (-x / h != divc(-x, h, y)) ++bad;
    }
    printf("%.9g %ld\n", h, bad);
  }
  return 0;
}
