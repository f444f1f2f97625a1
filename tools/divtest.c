#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static uint64_t s=88172645463325252ull; static inline uint64_t rnd(){s^=s<<13;s^=s>>7;s^=s<<17;return s;}
static inline double div_small(double num,int den,const double*rcp){double d=(double)den;double y=rcp[den];double q=num*y;double r=fma(-d,q,num);q=fma(r,y,q);r=fma(-d,q,num);return fma(r,y,q);}
static inline double div_one(double num,int den,const double*rcp){double d=(double)den;double y=rcp[den];double q=num*y;double r=fma(-d,q,num);return fma(r,y,q);}
/* divtest.c -- the table-based division of csrc/pava.cuh (div_small: RN(1/den) from a table, q0 = RN(num*y), two FMA
   residual corrections) against the true quotient, bit for bit: random mantissas over the guarded exponent window and
   constructions near rounding midpoints; den = 2..8192.  Also counts how often ONE correction would already do.
     gcc -O2 -march=native -ffp-contract=off divtest.c -lm && ./a.out [random cases (default 4e8)] */
int main(int argc,char**argv){static double rcp[8193];long NR=argc>1?atol(argv[1]):400000000L;for(int i=1;i<=8192;i++)rcp[i]=1.0/i;
 long bad=0,bad1=0,n=0;
 // random mantissas, exponents in the safe window
 for(long it=0;it<NR;it++){uint64_t m=rnd();int e=(int)(rnd()%1900)-900+1023; if(e<123)e=123; if(e>2023)e=2023; uint64_t bits=((uint64_t)(m&1)<<63)|((uint64_t)e<<52)|(m>>12);double a;memcpy(&a,&bits,8);int den=2+(int)(rnd()%8191);
  double t=a/den; double q=div_small(a,den,rcp); if(memcmp(&t,&q,8)){bad++;if(bad<5)printf("BAD %a / %d: %a vs %a\n",a,den,t,q);} double q1=div_one(a,den,rcp); if(memcmp(&t,&q1,8))bad1++; n++;}
 // near-midpoint constructions: a = den * (q + 0.5ulp) rounded
 for(long it=0;it<NR/2;it++){uint64_t m=rnd();uint64_t bits=((uint64_t)1023<<52)|(m>>12);double q;memcpy(&q,&bits,8);int den=2+(int)(rnd()%8191);
  double hq=q+ldexp(1.0,-53); /* rounds to q or next, use long double */ long double mid=(long double)q+ldexpl(1.0L,-53); long double al=mid*den; double a=(double)al; for(int dlt=-1;dlt<=1;dlt++){double aa=nextafter(a,dlt<0?-INFINITY:INFINITY); if(dlt==0)aa=a; double t=aa/den; double qq=div_small(aa,den,rcp); if(memcmp(&t,&qq,8)){bad++;if(bad<5)printf("BADM %a / %d\n",aa,den);} double q1=div_one(aa,den,rcp); if(memcmp(&t,&q1,8))bad1++; n++;} (void)hq;}
 printf("n=%ld bad(two-step)=%ld bad(one-step)=%ld\n",n,bad,bad1);return bad?1:0;}
