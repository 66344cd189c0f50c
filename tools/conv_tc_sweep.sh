# tcgen05 conv: accuracy / speed per drain granularity, then the pipeline's floor with stages switched off
for d in 3 1 9; do echo "== DRAIN $d"; PAIG_CONV_TC_DRAIN=$d timeout 300 python tools/conv_tc_probe.py big; done
echo "== small shapes, DRAIN 3"; timeout 300 python tools/conv_tc_probe.py
for g in 3 4 8 15; do echo "== DRAIN 3 DBG $g"; PAIG_CONV_TC_DBG=$g timeout 100 python tools/conv_tc_probe.py '[1000, 64, 64, 16, 1, 0]'; PAIG_CONV_TC_DBG=$g timeout 100 python tools/conv_tc_probe.py '[1000, 128, 128, 8, 1, 0]'; done
