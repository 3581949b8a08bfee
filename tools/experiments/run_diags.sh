#!/bin/bash
cp blurry_edges_b200/libblurry_edges_b200.so /tmp/lib_normal.so
echo "== normal"; python tools/prof_train.py 32 20 same | head -1
for n in 2 3 4; do cp /tmp/lib_diag$n.so blurry_edges_b200/libblurry_edges_b200.so; echo "== diag $n"; timeout 120 python tools/prof_train.py 32 20 same | head -1; done
cp /tmp/lib_normal.so blurry_edges_b200/libblurry_edges_b200.so
