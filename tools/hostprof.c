// tools/hostprof.c -- LD_PRELOAD sampling profiler (SIGPROF on process CPU time; program counter + 64 KB of stack per sample, read with
// process_vm_readv so that a short stack cannot fault); tools/hostprof_sym.py attributes each sample to the innermost hookable region found on
// its stack (scanning for return addresses: the reference is built without frame pointers).  gcc -O2 -shared -fPIC -o hostprof.so tools/hostprof.c;
// LD_PRELOAD=./hostprof.so PROF_OUT=/tmp/p oracle/_ref/turing_ref encode ...; python tools/hostprof_sym.py /tmp/p
#define _GNU_SOURCE
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/time.h>
#include <ucontext.h>
#include <unistd.h>
#include <sys/uio.h>
#define MAXS 6000
#define W 8192
static unsigned long *buf; static volatile long n;
static void handler(int sig, siginfo_t *si, void *ctx){ ucontext_t *uc=ctx; long i=__sync_fetch_and_add(&n,1); if(i>=MAXS) return; unsigned long *o=buf+i*(W+1); o[0]=uc->uc_mcontext.gregs[REG_RIP]; unsigned long *sp=(unsigned long*)uc->uc_mcontext.gregs[REG_RSP]; 
  /* copy stack words; stay within the page run that is surely mapped: stop at 8 KB boundary heuristics */
  struct iovec l={o+1,W*8}, r={sp,W*8}; process_vm_readv(getpid(),&l,1,&r,1,0); }
__attribute__((constructor)) static void init(void){ buf=calloc((size_t)MAXS*(W+1),sizeof(long)); struct sigaction sa; memset(&sa,0,sizeof sa); sa.sa_sigaction=handler; sa.sa_flags=SA_SIGINFO|SA_RESTART; sigaction(SIGPROF,&sa,0); struct itimerval it={{0,1000},{0,1000}}; setitimer(ITIMER_PROF,&it,0); }
__attribute__((destructor)) static void fini(void){ struct itimerval it={{0,0},{0,0}}; setitimer(ITIMER_PROF,&it,0); const char *o=getenv("PROF_OUT"); if(!o) o="/tmp/prof/out"; char p[512]; snprintf(p,512,"%s.stk",o); FILE*f=fopen(p,"wb"); long m=n<MAXS?n:MAXS; fwrite(buf,sizeof(long)*(W+1),m,f); fclose(f); snprintf(p,512,"%s.maps",o); FILE*g=fopen(p,"w"); FILE*mf=fopen("/proc/self/maps","r"); char line[1024]; while(fgets(line,1024,mf)) fputs(line,g); fclose(g); fclose(mf);}
