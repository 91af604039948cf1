# r4a: anchor segment (columns per k_anchor thread) against the strip length, full C4 map on one GPU
for st in 64 16; do for a in 16 8 4 2; do
  echo "strip $st anchor $a: $(MLM_MAP_STRIP=$st MLM_MAP_ANCHOR=$a python tools/profile_map.py 8192 5 2>&1 | sort -t: -k2 -n | sed -n 2p)"
done; done
