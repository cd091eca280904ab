#!/bin/bash
# A/B/n of prebuilt libb200fwd.so variants on the SAME box, interleaved (box-to-box spread is +-5 %, larger than most
# kernel tweaks).  Build the variants here (`python llm_np_cp_b200/build.py --force`), copy each to ab_<name>.so at the
# repo root (*.so files travel with gpurun), then:
#   gpurun -- 'tools/ab_libs.sh "python tools/gpu_diag.py attn_cta"'
cmd=${1:-"python tools/gpu_diag.py attn_cta"}
cp llm_np_cp_b200/libb200fwd.so /tmp/b2_keep.so
for rep in 1 2 3; do
  for f in ab_*.so; do
    cp "$f" llm_np_cp_b200/libb200fwd.so
    echo "== variant $f rep $rep"
    eval "$cmd" 2>&1 | grep "== attn\|tokens_per_sec" | cut -c1-220
  done
done
cp /tmp/b2_keep.so llm_np_cp_b200/libb200fwd.so
