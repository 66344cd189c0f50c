for t in "mnist_spring_color 16" "spring_color 100"; do
 echo "== $t: FMA convs, sgemm l1";            PAIG_NO_CONV_TC=1 PAIG_TC_MASK=6 python tools/logit_bias.py $t
 echo "== $t: FMA convs, drained tc l1";       PAIG_NO_CONV_TC=1 python tools/logit_bias.py $t
 echo "== $t: default (tc convs drain 3 + comp, drained tc l1)"; python tools/logit_bias.py $t
 echo "== $t: tc convs drain 3 no comp, sgemm l1"; PAIG_CONV_TC_NOCOMP=1 PAIG_TC_MASK=6 python tools/logit_bias.py $t
 echo "== $t: tc convs drain 3 + comp, sgemm l1"; PAIG_TC_MASK=6 python tools/logit_bias.py $t
 echo "== $t: tc convs drain 1 no comp, sgemm l1"; PAIG_CONV_TC_DRAIN=1 PAIG_CONV_TC_NOCOMP=1 PAIG_TC_MASK=6 python tools/logit_bias.py $t
 echo "== $t: tc convs drain 9 no comp, sgemm l1"; PAIG_CONV_TC_DRAIN=9 PAIG_CONV_TC_NOCOMP=1 PAIG_TC_MASK=6 python tools/logit_bias.py $t
done
