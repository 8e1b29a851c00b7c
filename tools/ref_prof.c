/* tools/ref_prof.c -- MEASUREMENT INFRASTRUCTURE: an LD_PRELOAD timer for the seams of `longcallD call` (oracle/_ref/longcallD_so, the
 * unmodified reference linked as a shared library).  Every wrapped function forwards to the next definition in the lookup order (the GPU
 * drop-in when it is preloaded after this library, else the reference) and adds its wall and thread-CPU seconds to a table printed at exit:
 *     LD_PRELOAD="tools/_build/libref_prof.so longcalld_b200/dropin/liblcd_dropin.so" oracle/_ref/longcallD_so call ...
 * Inclusive times, summed over the worker threads.  Not part of the product. */
#define _GNU_SOURCE
#include <dlfcn.h>
#include <stdio.h>
#include <time.h>
#include <pthread.h>
static double nowc(clockid_t c){struct timespec t;clock_gettime(c,&t);return t.tv_sec+1e-9*t.tv_nsec;}
#define now() nowc(CLOCK_MONOTONIC)
#define nowcpu() nowc(CLOCK_THREAD_CPUTIME_ID)
#define N 32
static double acc[N], accc[N]; static unsigned long cnt[N]; static const char *nm[N]; static pthread_mutex_t mu=PTHREAD_MUTEX_INITIALIZER;
static void add(int i,const char*n,double dt,double dc){pthread_mutex_lock(&mu);acc[i]+=dt;accc[i]+=dc;cnt[i]++;nm[i]=n;pthread_mutex_unlock(&mu);}
#define WRAP(name, idx) void name(void){ static void(*o)(void); if(!o) o=(void(*)(void))dlsym(RTLD_NEXT,#name); void *args=__builtin_apply_args(); double t0=now(),c0=nowcpu(); void *ret=__builtin_apply(o,args,512); add(idx,#name,now()-t0,nowcpu()-c0); __builtin_return(ret);}
WRAP(collect_var_main,0)
WRAP(collect_digars_from_bam,1)
WRAP(collect_all_cand_var_sites,2)
WRAP(collect_cand_vars,3)
WRAP(pre_process_noisy_regs,4)
WRAP(classify_cand_vars,5)
WRAP(collect_read_var_profile,6)
WRAP(assign_hap_based_on_germline_het_vars_kmeans,7)
WRAP(collect_noisy_vars1,8)
WRAP(collect_noisy_reg_aln_strs,9)
WRAP(collect_noisy_read_info,10)
int abpoa_partial_aln_msa_cons(const void *opt, void *ab, int sampling_reads, int n_reads, int *read_ids, void *a, void *b, int *read_lens, int *fc, void *c, int max_n_cons, void *d, void *e, void *f, void *g, void *h, void *i){
 static int (*o)(const void*,void*,int,int,int*,void*,void*,int*,int*,void*,int,void*,void*,void*,void*,void*,void*); if(!o) o=dlsym(RTLD_NEXT,"abpoa_partial_aln_msa_cons");
 int ok = ab==0 && !sampling_reads && max_n_cons==1 && n_reads>=1 && (fc[0]&12)==12; int part=0;
 for(int k=0;k<n_reads;++k){int cv=fc[k]; int both=(cv&12)==12, lc=(cv&8)&&!(cv&4), rc=!(cv&8)&&(cv&4); if(!(both || (lc&&(cv&1)) || (rc&&(cv&2)))) {ok=0;part=1;}}
 double t0=now(),c0=nowcpu(); int r=o(opt,ab,sampling_reads,n_reads,read_ids,a,b,read_lens,fc,c,max_n_cons,d,e,f,g,h,i); double dt=now()-t0,dc=nowcpu()-c0;
 add(11,"abpoa_partial_aln_msa_cons",dt,dc); if(ok) add(29,"  poa: full-cover, 1 consensus",dt,dc); else { add(30,"  poa: partial-cover / sampled / 2 consensus",dt,dc); if(sampling_reads) add(31,"  poa:   of those, sampled",dt,dc);} return r;}
WRAP(abpoa_aln_msa_cons,12)
WRAP(wfa_end2end_aln,13)
WRAP(edlib_xgaps,14)
WRAP(edlib_end2end_aln,15)
WRAP(edlib_infix_aln,16)
WRAP(make_vars_from_msa_cons_aln,17)
WRAP(wfa_collect_noisy_aln_str_with_ps_hap,18)
WRAP(wfa_collect_noisy_aln_str_no_ps_hap,19)
WRAP(collect_ref_seq_bam_main,20)
WRAP(stitch_var_main,21)
WRAP(make_var_main,22)
WRAP(collect_somatic_var,23)
WRAP(sort_noisy_region_reads,24)
WRAP(make_cons_read_aln_str,25)
WRAP(wfa_collect_aln_str,26)
WRAP(bam_chunk_free,27)
WRAP(post_process_noisy_regs,28)
__attribute__((destructor)) static void rep(void){ for(int i=0;i<N;++i) if(nm[i]) fprintf(stderr,"[prof] %-50s wall %9.3f s  cpu %9.3f s  %8lu calls\n",nm[i],acc[i],accc[i],cnt[i]); }
