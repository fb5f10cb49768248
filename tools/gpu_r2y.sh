for wl in c2 c3 c5; do
st=50; [ $wl = c5 ] && st=10
for w in "1,0.95,0.90,0.865" "1,0.96,0.92,0.88" "1,0.97,0.93,0.885" "1,0.98,0.93,0.86" "1,1,0.92,0.85" "1,0.97,0.94,0.91"; do
  echo "$wl W=$w $(PBRT_B200_RANK_W=$w python tools/pdl_probe.py $wl $st 2>/dev/null | tail -1)"
done; done
