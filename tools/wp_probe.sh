# whole-program probe: reference vs GPU drop-in at several worker-thread counts on one synthetic BAM (scratch tool)
mkdir -p /tmp/wp && cd /tmp/wp && [ -f s50.bam.bai ] || /root/repo/tools/_build/synth_bam s50 ${MB:-50} hifi 11 >/dev/null 2>&1
R=/root/repo
P=$R/tools/_build/libref_prof.so
D=$R/longcalld_b200/dropin/liblcd_dropin.so
export LCD_DROPIN_VERBOSE=1
$R/oracle/_ref/longcallD_ref call --hifi -t 16 s50.fa s50.bam 2>ref.err | grep -v '^#' | md5sum; grep Real ref.err
for t in ${THREADS:-16 32 64 100}; do
  for st in ${STAGES:-all engines}; do
    echo "== -t $t stages $st"
    LCD_DROPIN_STAGES=$st LD_PRELOAD="$D" $R/oracle/_ref/longcallD_so call --hifi -t $t s50.fa s50.bam 2>gpu.err | grep -v '^#' | md5sum; grep "Real\|dropin" gpu.err
  done
done
