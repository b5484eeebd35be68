#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
static inline float u2f(uint32_t u){float f; memcpy(&f,&u,4); return f;}
static inline uint32_t f2u(float f){uint32_t u; memcpy(&u,&f,4); return u;}
int main(int argc,char**argv){
  int wmax = atoi(argv[1]); int variant = atoi(argv[2]);
  long long bad=0; int firstbad=1<<30;
  #pragma omp parallel for schedule(dynamic,64) reduction(+:bad) reduction(min:firstbad)
  for(int w=0; w<wmax; ++w){
    float wf=(float)w, w1=wf+1.0f;
    float rc = 1.0f/w1;
    for(int dr=-1; dr<=1; ++dr){
      float r0 = u2f(f2u(rc)+dr);
      float r = (variant&1) ? r0 : fmaf(r0, fmaf(r0,-w1,1.0f), r0);
      for(int o=0;o<256;++o) for(int n=0;n<256;++n){
        volatile float p = wf*(float)o; volatile float a = p + (float)n; volatile float Q = a / w1;
        float yr = fminf(255.0f, rintf(Q));
        float op = u2f(0x4B000000u | (uint32_t)o), np_ = u2f(0x4B000000u | (uint32_t)n);
        float d = np_ - op;
        float q0 = d*r;
        float q = (variant&2) ? q0 : fmaf(r, fmaf(q0,-w1,d), q0);
        float y = (variant&4) ? fmaf(d, r, op) : q + op;
        uint32_t yb = f2u(y) & 0xFFu;
        if (yb != (uint32_t)yr) { bad++; if (w<firstbad) firstbad=w; }
      }
    }
  }
  printf("variant=%d wmax=%d mismatches=%lld first bad w=%d\n", variant, wmax, bad, firstbad);
  return 0;
}
