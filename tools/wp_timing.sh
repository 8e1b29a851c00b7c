mkdir -p /tmp/wp && cd /tmp/wp && [ -f s50.bam.bai ] || /root/repo/tools/_build/synth_bam s50 50 hifi 11 >/dev/null 2>&1
R=/root/repo; D=$R/longcalld_b200/dropin/liblcd_dropin.so
mkdir -p /tmp/vt && cp $R/build/vt/liblcd_gpu.so /tmp/vt/
LCD_POA_TIMING_PRINT=1 LD_LIBRARY_PATH=/tmp/vt LCD_DROPIN_STAGES=engines LD_PRELOAD="/tmp/vt/liblcd_gpu.so $D" $R/oracle/_ref/longcallD_so call --hifi -t 16 s50.fa s50.bam 2>gpu.err >/dev/null
grep -c "batch Mcycles" gpu.err
grep "poa timing\] #" gpu.err | awk '{for(i=1;i<=NF;i++) if($i=="dp") print $(i+1), $0}' | sort -rn | head -12 | cut -c1-260
grep "batch Mcycles" gpu.err | awk '{dp+=$5; seg+=$7; gen+=$13; bt+=$19; add+=$21; aft+=$23; fin+=$28; sub+=$32} END {print "total Mcycles dp",dp,"seg",seg,"gen",gen,"bt",bt,"add",add,"after",aft,"fin",fin,"sub",sub}'
grep "Real" gpu.err
