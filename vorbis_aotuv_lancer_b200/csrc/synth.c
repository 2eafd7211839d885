/* Deterministic synthetic PCM of the shape SURVEY.md section 8(d) names: two
 * swept/steady tones, a high tone, low-passed noise, a noise floor and a
 * decaying click every 0.5 s (so all four block types and the aoTuV M3 impulse
 * path are exercised).  Host-only, no reference code involved. */
#include <math.h>
#include <stdint.h>
#include "vorbis_b200.h"

static inline uint64_t xs64(uint64_t *s){
  uint64_t x=*s; x^=x<<13; x^=x>>7; x^=x<<17; *s=x; return x;
}
static inline double uni(uint64_t *s){ /* U(-1,1) */
  return (double)(xs64(s)>>11)*(2.0/9007199254740992.0)-1.0;
}
uint64_t vb200_splitmix64(uint64_t x){
  x+=0x9E3779B97F4A7C15ull;
  x=(x^(x>>30))*0xBF58476D1CE4E5B9ull;
  x=(x^(x>>27))*0x94D049BB133111EBull;
  return x^(x>>31);
}

int vb200_synth_pcm(uint64_t seed,int channels,long rate,long nsamples,float *out,long stride){
  double lp[VB200_MAX_CHANNELS];
  const double twopi=6.283185307179586476925286766559;
  long i; int c;
  if(channels<1||channels>VB200_MAX_CHANNELS||!out)return VB200_EINVAL;
  if(seed==0)seed=88172645463325252ull;
  for(c=0;c<channels;c++)lp[c]=0.;
  for(i=0;i<nsamples;i++){
    double t=(double)i/(double)rate;
    double tm=fmod(t,0.5);
    double env=0.5*exp(-40.*tm);
    for(c=0;c<channels;c++){
      double n=uni(&seed), u=uni(&seed), v;
      lp[c]=0.98*lp[c]+0.02*n;
      v=0.25*sin(twopi*(220.+30.*sin(0.7*t))*t+0.3*c)
       +0.15*sin(twopi*1760.*t+0.6*c)
       +0.08*sin(twopi*6500.*t)
       +0.6*lp[c]+0.02*n
       +env*u*((c&1)?0.7:1.0);
      if(v>0.99)v=0.99;
      if(v<-0.99)v=-0.99;
      out[(long)c*stride+i]=(float)v;
    }
  }
  return 0;
}
