# whole-program sweep of the drop-in's knobs on one synthetic BAM (scratch tool): "inflight chunks-per-thread pool-GB linger-us" per line, 2 runs each
mkdir -p /tmp/wp && cd /tmp/wp && [ -f s50.bam.bai ] || /root/repo/tools/_build/synth_bam s50 50 hifi 11 >/dev/null 2>&1
R=/root/repo; D=$R/longcalld_b200/dropin/liblcd_dropin.so
for i in 1 2; do $R/oracle/_ref/longcallD_ref call --hifi -t 16 s50.fa s50.bam 2>ref.err >/dev/null; grep -o "Real time: [0-9.]*" ref.err; done
for cfg in "3 2 48 300" "3 3 48 300" "3 4 48 300" "4 3 64 300" "2 3 48 300" "3 3 48 100" "3 3 48 1000"; do
  set -- $cfg
  printf "inflight %s cpt %s pool %s linger %s: " $1 $2 $3 $4
  for i in 1 2 3; do LCD_DROPIN_INFLIGHT=$1 LCD_DROPIN_CHUNKS_PER_THREAD=$2 LCD_DROPIN_POOL_GB=$3 LCD_DROPIN_LINGER_US=$4 LCD_DROPIN_STAGES=engines LD_PRELOAD="$D" $R/oracle/_ref/longcallD_so call --hifi -t 16 s50.fa s50.bam 2>gpu.err >/dev/null; grep -o "Real time: [0-9.]*" gpu.err | tr '\n' ' '; done; echo
done
