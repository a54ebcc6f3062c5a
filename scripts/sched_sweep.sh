export SM_PROBE_OVERLAP_ONLY=1
for at in scan sm; do
 for c in 48 64 72 88 120; do
  echo "== coop ctas=$c aux_at=$at"; SMG_AUX_AT=$at SMG_SM_CTAS=$c python scripts/sm_probe.py 100000 256 50 60 coop 2>&1 | head -2
 done
 echo "== cluster aux_at=$at"; SMG_AUX_AT=$at python scripts/sm_probe.py 100000 256 50 60 cluster 2>&1 | head -2
done
