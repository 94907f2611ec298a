// exhaustive check: q = fma(fma(-q0, c, x), rc, q0), q0 = x*rc, rc = RN(1/c)  ==  x / c  for all floats x in [0, c]
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
static const float CS[] = {10.f, 60.f, 1200.f, 120.f, 40.f};
typedef struct { float c; uint32_t lo, hi; long bad; } job;
static void *run(void *a){ job *j=a; float c=j->c, rc=1.0f/c; long bad=0;
  for(uint32_t u=j->lo; u<j->hi; ++u){ float x; memcpy(&x,&u,4); float q0=x*rc; float r=fmaf(-q0,c,x); float q=fmaf(r,rc,q0); float w=x/c; if(memcmp(&q,&w,4)) bad++; }
  j->bad=bad; return 0; }
int main(){ for(int k=0;k<5;k++){ float c=CS[k]; uint32_t top; memcpy(&top,&c,4); top+=1; float lo0=1e-30f; uint32_t bot; memcpy(&bot,&lo0,4); int T=8; pthread_t th[8]; job jb[8];
    for(int t=0;t<T;t++){ jb[t].c=c; jb[t].lo=bot+(uint32_t)((uint64_t)(top-bot)*t/T); jb[t].hi=bot+(uint32_t)((uint64_t)(top-bot)*(t+1)/T); pthread_create(&th[t],0,run,&jb[t]); }
    long bad=0; for(int t=0;t<T;t++){ pthread_join(th[t],0); bad+=jb[t].bad; }
    printf("c=%g: %u floats in [0,c], mismatches %ld\n", c, top, bad); }
  return 0; }
