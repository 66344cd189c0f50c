# tcgen05 conv: accuracy / speed per drain granularity, with and without the truncation-bias compensation
for d in 3 1 9; do echo "== DRAIN $d + compensation"; PAIG_CONV_TC_DRAIN=$d timeout 300 python tools/conv_tc_probe.py big; done
echo "== DRAIN 3, no compensation"; PAIG_CONV_TC_NOCOMP=1 timeout 300 python tools/conv_tc_probe.py big
