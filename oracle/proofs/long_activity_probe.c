/* long_activity_probe.c -- TEST INFRASTRUCTURE (design probe, round 2): how often the long peak detector is stepped, holds a
 * peak above its threshold, or emits, on synthetic reads and on a dump of sp1_dna.blow5 (profiles/r02_filter_probe.txt).
 * build: gcc -O2 -std=c99 -ffp-contract=off -I../.. long_activity_probe.c ../sigtk_oracle.c -lm */
#define _POSIX_C_SOURCE 200809L
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "../sigtk_oracle.h"
static uint64_t rng_s = 88172645463325252ull;
static double urand(void){rng_s ^= rng_s << 13; rng_s ^= rng_s >> 7; rng_s ^= rng_s << 17;return (double)(rng_s >> 11) * (1.0 / 9007199254740992.0);}
static double nrand(void){double u1=urand(),u2=urand(); if(u1<1e-300)u1=1e-300; return sqrt(-2.0*log(u1))*cos(6.283185307179586*u2);}
int main(int argc,char**argv){
  int rna=atoi(argv[1]); const char*file=argc>2?argv[2]:NULL; double noise=argc>3?atof(argv[3]):2.0;
  orc_params_t P; orc_params(rna,&P);
  FILE*f=file&&strcmp(file,"-")?fopen(file,"rb"):NULL; int64_t n_reads=20; if(f) fread(&n_reads,8,1,f);
  uint64_t steps=0,l_unm=0,l_hot=0,sup=0,sup8=0,emits=0,eml=0,blocks=0,bh=0,bs=0,chunks=0,ch_hot=0,ch_sup=0;
  for(int r=0;r<n_reads;r++){
    int64_t n=40000; double dig=8192.0,off=r%53,range=1402.882324;
    if(f){ fread(&n,8,1,f); fread(&dig,8,1,f); fread(&off,8,1,f); fread(&range,8,1,f);} 
    int16_t*raw=malloc(n*2); float*pa=malloc(n*4); double*S=malloc((n+1)*8),*Q=malloc((n+1)*8); float*t1=malloc(n*4),*t2=malloc(n*4);
    if(f) fread(raw,2,n,f); else { double level=60+60*urand(), pch=rna?0.025:0.1; for(int64_t i=0;i<n;i++){ if(urand()<pch) level=60+60*urand(); raw[i]=(int16_t)rint((level+noise*nrand())*(dig/range)-off);} }
    orc_pa(raw,n,dig,off,range,pa); orc_prefix(pa,n,S,Q); orc_tstat(S,Q,n,P.w_short,t1); orc_tstat(S,Q,n,P.w_long,t2);
    orc_det_t s,l; orc_det_init(&s,&l); uint64_t pk[4]; int hb=0,sb=0,hc=0,sc=0;
    for(int64_t i=0;i<n;i++){
      orc_det_t s0=s,l0=l;
      uint64_t np=orc_detect(t1,t2,i,i+1,&P,&s,&l,pk,4);
      emits+=np; int short_emit=(s0.peak_pos>=0&&s.peak_pos<0); if(np>short_emit) eml++;
      steps++;
      /* was the long detector stepped at i? it is stepped iff masked_to (after the short step) < i */
      int stepped = (l.masked_to < (uint64_t)i);
      if(stepped){ l_unm++; if(t2[i]>P.thr_long*0.99f){sup++; sb=1; sc=1;} if(l.peak_pos>=0&&l.peak_value>P.thr_long){l_hot++; hb=1; hc=1;} }
      (void)l0;
      if((i&7)==7){blocks++; bh+=hb; bs+=sb; hb=sb=0;}
      if((i&1023)==1023){chunks++; ch_hot+=hc; ch_sup+=sc; hc=sc=0;}
    }
    free(raw);free(pa);free(S);free(Q);free(t1);free(t2);
  }
  printf("rna=%d file=%s steps=%lu: long stepped %.3f, hot %.2e/step, superset(t2>0.99thr & stepped) %.2e/step; blocks hot %.2e sup %.2e; 1024-chunks hot %.3f sup %.3f; emits %.4f/sample of which long %.2e/sample\n",rna,file?file:"synth",(unsigned long)steps,(double)l_unm/steps,(double)l_hot/steps,(double)sup/steps,(double)bh/blocks,(double)bs/blocks,(double)ch_hot/(chunks?chunks:1),(double)ch_sup/(chunks?chunks:1),(double)emits/steps,(double)eml/steps);
  return 0; }
