for wl in c2 c3; do
for w in "1,0.95,0.90,0.865" "1,0.97,0.94,0.91" "1,0.93,0.87,0.82" "1,0.9,0.82,0.76" "1,0.98,0.93,0.86" "1,1,0.92,0.85"; do
  echo "$wl W=$w $(PBRT_B200_RANK_W=$w python tools/pdl_probe.py $wl 50 2>/dev/null | tail -1)"
done; done
echo "c2 halo"; for c in 0.6 2.0; do echo "C=$c $(PBRT_B200_HALO_C=$c python tools/pdl_probe.py c2 50 2>/dev/null | tail -1)"; done
