#include <pthread.h>
#include <stdio.h>
#include <time.h>
#include <stdlib.h>
static void* w(void* a){ volatile unsigned long x=0; for(unsigned long i=0;i<300000000ul;i++) x+=i; return 0;}
int main(int c,char**v){int T=atoi(v[1]); pthread_t t[256]; struct timespec a,b; clock_gettime(CLOCK_MONOTONIC,&a);
for(int i=0;i<T;i++)pthread_create(&t[i],0,w,0); for(int i=0;i<T;i++)pthread_join(t[i],0); clock_gettime(CLOCK_MONOTONIC,&b);
printf("%d threads %.3f s\n",T,(b.tv_sec-a.tv_sec)+(b.tv_nsec-a.tv_nsec)*1e-9);}
