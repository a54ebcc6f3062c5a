# gang size of the cooperative split-merge kernel x where the next pass's aux columns are evaluated (metric shape)
export SM_PROBE_OVERLAP_ONLY=1
for c in ${CTAS:-40 48 56 64 72 88}; do
  echo "== coop ctas=$c"; SMG_SM_CTAS=$c python scripts/sm_probe.py 100000 256 50 60 coop 2>&1 | head -2
done
