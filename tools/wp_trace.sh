# whole-program probe of the drop-in's knobs on one synthetic BAM (scratch tool): repeated runs with a timeline each
mkdir -p /tmp/wp && cd /tmp/wp && [ -f s50.bam.bai ] || /root/repo/tools/_build/synth_bam s50 ${MB:-50} hifi 11 >/dev/null 2>&1
R=/root/repo; D=$R/longcalld_b200/dropin/liblcd_dropin.so
export LCD_DROPIN_VERBOSE=1
$R/oracle/_ref/longcallD_ref call --hifi -t 16 s50.fa s50.bam 2>ref.err | grep -v '^#' | md5sum; grep Real ref.err
for rep in ${REPS:-1 2 3 4}; do
  echo "== run $rep"
  LCD_DROPIN_TRACE=$R/gpurun_out/trace_rep$rep.tsv LCD_DROPIN_STAGES=${STAGES:-engines} LD_PRELOAD="$D" $R/oracle/_ref/longcallD_so call --hifi -t ${THREADS:-16} s50.fa s50.bam 2>gpu.err | grep -v '^#' | md5sum; grep "Real\|dropin" gpu.err
done
nvidia-smi --query-gpu=clocks.sm,clocks_throttle_reasons.active,temperature.gpu,power.draw --format=csv
uptime
