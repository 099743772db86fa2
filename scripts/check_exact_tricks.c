// Developer check (CPU): exhaustive verification of the exact-arithmetic shortcuts of kernels.cu.
// gcc -O2 -ffp-contract=off -mfma -o /tmp/check scripts/check_exact_tricks.c -lm && /tmp/check
// exhaustive check: Markstein division r/R with R = 0.04f, and threshold tricks
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline float u2f(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t f2u(float f){uint32_t u; memcpy(&u,&f,4); return u;}
int main(){
  const float R = (float)(2.*(double)0.02f);
  const float y = (float)(1.0/(double)R);
  unsigned long bad=0, n=0;
  for (uint32_t u = f2u(1e-30f); u <= f2u(1e30f); ++u){
    float r = u2f(u);
    float q0 = r*y;
    float rem = fmaf(-q0, R, r);
    float q1 = fmaf(rem, y, q0);
    float ex = r / R;
    if (q1 != ex) { if (bad<5) printf("bad r=%a q1=%a ex=%a\n", r, q1, ex); ++bad; }
    ++n;
  }
  printf("markstein: tested %lu bad %lu  R=%a y=%a\n", n, bad, R, y);
  return 0;
}
