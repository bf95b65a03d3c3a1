/* Exhaustive-ish check of the quotient used by fd_div_by (lpm-c_b200/csrc/lpmb_stiffness.cu): x / y from rcp = RN(1 / y) and two
 * FMA-residual corrections equals the IEEE quotient.  usage: check_exact_division [samples per divisor = 60000000]
 * build: gcc -O2 -march=native -ffp-contract=off check_exact_division.c -o check_exact_division -lm   (exit code 1 on a mismatch) */
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <stdint.h>
#include <string.h>
static inline double divr(double x, double y, double r){
    double q0 = x*r;
    double e0 = fma(-y,q0,x);
    double q1 = fma(e0,r,q0);
    double e1 = fma(-y,q1,x);
    return fma(e1,r,q1);
}
static uint64_t s=88172645463325252ull;
static inline uint64_t rnd(){ s^=s<<13; s^=s>>7; s^=s<<17; return s;}
int main(int argc, char **argv){
    const long NS = argc > 1 ? atol(argv[1]) : 60000000;
    double ys[]={1e-6,0.25,0.3,0.1,0.05,1e-6*0.3, 0.7071067811865476, 3.0, 0.123456789, 1.9999999999999998, 1.0000000000000002};
    long bad=0, tot=0;
    for(int k=0;k<11;k++){
        double y=ys[k], r=1.0/y;
        for(long i=0;i<NS;i++){
            uint64_t b=rnd();
            // random significand, exponent in [-300,300]
            uint64_t mant=b&0xFFFFFFFFFFFFFull; int e=(int)((b>>52)%600)-300; uint64_t sign=(b>>63);
            uint64_t bits=(sign<<63)|((uint64_t)(e+1023)<<52)|mant;
            double x; memcpy(&x,&bits,8);
            double a=x/y, c=divr(x,y,r);
            tot++;
            if(a!=c){ bad++; if(bad<10) printf("y=%.17g x=%.17g %.17g vs %.17g\n",y,x,a,c);}
        }
    }
    // random y too
    for(long i=0;i<3*NS;i++){
        uint64_t b=rnd(), b2=rnd();
        uint64_t bits=(b&0x800FFFFFFFFFFFFFull)|((uint64_t)(1023+ (int)((b>>52)%200)-100)<<52);
        uint64_t bits2=(b2&0x000FFFFFFFFFFFFFull)|((uint64_t)(1023+ (int)((b2>>52)%60)-30)<<52);
        double x,y; memcpy(&x,&bits,8); memcpy(&y,&bits2,8);
        double r=1.0/y; double a=x/y,c=divr(x,y,r); tot++;
        if(a!=c){bad++; if(bad<20) printf("y=%.17g x=%.17g %.17g vs %.17g\n",y,x,a,c);}
    }
    printf("bad %ld of %ld\n",bad,tot);
    return bad ? 1 : 0;
}
