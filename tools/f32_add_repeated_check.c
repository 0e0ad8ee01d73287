// validation of the closed-form replay of k rounded float additions
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static float replay(float v,float d,int k){ volatile float x=v; for(int i=0;i<k;++i){ x = x + d; } return x; }
static int iters_total=0;
static float jump(float v,float d,int k)
{
	while (k > 0)
	{
		iters_total++;
		// three real steps (the third result starts the constant-step regime, see the tie analysis)
		volatile float t;
		t = v + d; float v1 = t; if (--k == 0) return v1;
		t = v1 + d; float v2 = t; if (--k == 0) return v2;
		t = v2 + d; float v3 = t; if (--k == 0) return v3;
		const uint32_t b1=f2u(v1), b2 = f2u(v2), b3 = f2u(v3);
		const uint32_t e = (b3 >> 23) & 0xFF;
		v = v3;
		if ((b1 >> 23) == (b2 >> 23) && (b2 >> 23) == (b3 >> 23) && e != 0 && e != 0xFF) // same sign, same binade, normal
		{
			const int32_t q = (int32_t)(b3 & 0x7FFFFF) - (int32_t)(b2 & 0x7FFFFF); // step of the magnitude, in ulps
			const int32_t m3 = (int32_t)(b3 & 0x7FFFFF);
			int32_t n; // steps that keep the magnitude strictly inside the binade: mantissa in [1, 2^23 - 1]
			if (q > 0) n = (0x7FFFFF - m3) / q;
			else if (q < 0) n = (m3 - 1) / (-q);
			else n = k;
			if (n > k) n = k;
			if (n > 0) { v = u2f(b3 + (uint32_t)(n * q)); k -= n; }
		}
	}
	return v;
}
static uint64_t s=88172645463325252ull; static uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s; }
int main(int argc,char**argv)
{
	long bad=0, n=atol(argv[1]); if(argc>2) s^=strtoull(argv[2],0,10)*0x9E3779B97F4A7C15ull;
	for(long i=0;i<n;++i)
	{
		float v,d; int k = (int)(rnd()%1100);
		int mode = rnd()%8;
		if(mode==0){ v=u2f((uint32_t)rnd()); d=u2f((uint32_t)rnd()); }
		else if(mode==1){ v = (float)(rnd()%100000)/977.0f; d = ((float)(rnd()%2000)-1000.0f)/(float)(1+rnd()%100000); }
		else if(mode==2){ // ties: d = (q+0.5) ulp(v)
			v = u2f(0x3F800000u + (uint32_t)(rnd()%0x7FFFFF)); float u=ldexpf(1.0f,-23); d = ((float)(rnd()%9)-4.0f+0.5f)*u; }
		else if(mode==3){ v = u2f(0x3F800000u + (uint32_t)(rnd()%0x7FFFFF)); float u=ldexpf(1.0f,-24); d = ((float)(rnd()%17)-8.0f+0.5f)*u; } // lower binade ties
		else if(mode==4){ v = ldexpf(1.0f,(int)(rnd()%10)) ; d = -ldexpf((float)(rnd()%1000),-30+(int)(rnd()%20)); } // starting at boundary going down
		else if(mode==5){ v = (float)(rnd()%1000)*1e-3f; d = (float)(rnd()%1000)*1e-6f - 5e-4f; } // crossing zero
		else if(mode==6){ v = u2f((uint32_t)(rnd()%0x01000000)); d = u2f((uint32_t)(rnd()%0x01000000)) * ((rnd()&1)?1.f:-1.f); } // denormals
		else { v = 1.0f/(1.0f+(float)(rnd()%4096)); d = v * ((float)(rnd()%2001)-1000.0f)*1e-5f; } // inv_z like
		if (v!=v || d!=d) continue;
		float a=replay(v,d,k), b=jump(v,d,k);
		if (f2u(a)!=f2u(b) && !(a!=a && b!=b)) { if(bad<10) printf("BAD mode %d v=%a d=%a k=%d replay=%a jump=%a\n",mode,v,d,k,a,b); bad++; }
	}
	printf("%ld cases, %ld bad, avg loop iterations %.2f\n", n, bad, (double)iters_total/n);
	return bad!=0;
}
