for l in 1 2 3 5; do
 for cfg in "2 222" "2 296" "2 592" "2 1184" "2 2368" "1 296" "1 592" "1 1184"; do
  set -- $cfg
  echo "layer $l persm=$1 ctas=$2: $(GEECO_TC_PERSM=$1 GEECO_TC_CTAS=$2 python tools/prof_conv.py --layer $l --what fwd | grep conv)"
 done
done
