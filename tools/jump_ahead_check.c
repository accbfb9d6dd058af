/* Evidence for kfb_integrate2.cuh `jump_ahead`: x <- fl(x + d) applied n times in fp32 (round to nearest even) computed in
 * O(binades crossed) instead of n additions, bit for bit.  Inside one binade and sign every correctly rounded addition of the
 * same d moves x by the SAME whole number of ulps (ties included once the parity has settled, i.e. from the second step in
 * the binade on), so after two real steps in a binade the remaining ones are one integer multiply-add on the bit pattern;
 * the step that leaves the binade is taken for real.  This program checks the routine against the plain loop on millions of
 * random, KinectFusion-like and tie-prone (x, d, n):
 *     gcc -O2 -ffp-contract=off tools/jump_ahead_check.c -lm -o /tmp/jump && /tmp/jump      (~1 min; exit status 0 = no mismatch)
 */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>
#include <math.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static long n_real=0,n_jump=0;
float jump(float x, float d, int n){
	volatile float vx;
	int k=n;
	while(k>0){
		float a=x+d; n_real++; --k; if(k==0){x=a;break;}
		float b=a+d; n_real++; --k;
		const uint32_t ux=f2u(x),ua=f2u(a),ub=f2u(b);
		const uint32_t ex=ux>>23,ea=ua>>23,eb=ub>>23;          // sign + exponent
		if(k>0 && ex==ea && ea==eb && (ea&0xff)!=0 && (ea&0xff)!=0xff){
			const int32_t S=(int32_t)(ub-ua);                    // ulps per step in magnitude space (same sign, same exponent)
			if(S==0){ x=b; k=0; break; }                         // stagnation: every further step returns b
			const uint32_t M=ub&0x7fffffffu, lo=(ea&0xff)<<23, hi=lo+0x7fffffu;
			uint32_t room = S>0 ? hi-M : (M==lo?0u:M-lo-1u); uint32_t m=(uint32_t)(((float)room/(float)abs(S))*0.9999f);   /* decreasing: stay STRICTLY above the binade floor (a sum just below it rounds on the finer grid) */
			if (M==lo && S<0) m=0;
			if(m>(uint32_t)k) m=k;
			b=u2f(ub+(uint32_t)((int32_t)m*S)); k-=m; n_jump+=m;
		}
		x=b;
	}
	(void)vx;
	return x;
}
float brute(float x,float d,int n){ for(int i=0;i<n;i++){ volatile float t=x+d; x=t;} return x; }
int main(){
	srand(7); long bad=0,tot=0;
	for(int t=0;t<6000000;t++){
		float x,d; int n;
		int mode=rand()%6;
		if(mode==0){ x=((rand()/(float)RAND_MAX)-0.5f)*10.f; d=((rand()/(float)RAND_MAX)-0.5f)*0.02f; }
		else if(mode==1){ x=((rand()/(float)RAND_MAX)-0.5f)*4000.f; d=((rand()/(float)RAND_MAX)-0.5f)*10.f; }
		else if(mode==2){ x=u2f((uint32_t)rand()<<1 ^ (uint32_t)rand()); d=u2f((uint32_t)rand()<<1 ^ (uint32_t)rand()); if(!(fabsf(x)<1e30f)||!(fabsf(d)<1e30f)) continue; }
		else if(mode==4){ float dz=4.8f/(float)(256<<(rand()%4)); float r=((rand()/(float)RAND_MAX)-0.5f)*0.2f; x=((rand()/(float)RAND_MAX)-0.5f)*6.f; d=dz*(rand()%2? 1.f: r); if(rand()%3==0) d=-d; }
		else if(mode==5){ float dz=4.8f/(float)(256<<(rand()%4)); x=((rand()/(float)RAND_MAX)-0.5f)*3000.f; d=dz*(481.2f*((rand()/(float)RAND_MAX)-0.5f)*0.2f+320.f*(0.9f+0.1f*(rand()/(float)RAND_MAX))); if(rand()%3==0) d=-d; }
		else { x=((rand()/(float)RAND_MAX)-0.5f)*5.f; d=ldexpf(((rand()%2000)-1000)/1024.f, -(rand()%24)); }   // ties likely
		n=rand()%2100;
		float j=jump(x,d,n), b=brute(x,d,n); tot++;
		if(f2u(j)!=f2u(b) && !(j!=j && b!=b)){ if(bad<10) printf("MISMATCH x=%a d=%a n=%d jump=%a brute=%a\n",x,d,n,j,b); bad++; }
	}
	printf("cases %ld mismatches %ld  real adds %ld jumped steps %ld\n",tot,bad,n_real,n_jump);
	return bad!=0;
}
