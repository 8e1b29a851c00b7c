mkdir -p /tmp/wp && cd /tmp/wp && [ -f s50.bam.bai ] || /root/repo/tools/_build/synth_bam s50 ${MB:-50} hifi 11 >/dev/null 2>&1
R=/root/repo; D=$R/longcalld_b200/dropin/liblcd_dropin.so
export LCD_DROPIN_VERBOSE=1
for f in ${INFLIGHT:-1 3}; do for t in ${THREADS:-16 32}; do
  echo "== inflight $f threads $t"
  LCD_DROPIN_INFLIGHT=$f LCD_DROPIN_TRACE=$R/gpurun_out/trace_f${f}_t$t.tsv LCD_DROPIN_STAGES=${STAGES:-engines} LD_PRELOAD="$D" $R/oracle/_ref/longcallD_so call --hifi -t $t s50.fa s50.bam 2>gpu.err | grep -v '^#' | md5sum; grep "Real\|dropin" gpu.err
done; done
