#!/bin/bash
# Scratch: A/B of build variants (tools/build_variant.py) on the GPU box.  Each argument is "variant[:ENV=VAL,ENV=VAL[:...]]"
# ("default" = the library as built); every run times the proxy and DarkCornell at 16 spp with stage times (tools/gpu_sweep.py).
# usage: bash tools/gpu_variant_sweep.sh OUTFILE variant[:settings] ...
out=$1; shift
: > $out
for spec in "$@"; do
    IFS=: read -r v s1 s2 s3 <<< "$spec"
    lib=""; [ "$v" != default ] && lib=$PWD/rust-path-tracer_b200/_build/variants/librpt_$v.so
    RPT_B200_LIBRARY=$lib timeout 300 python tools/gpu_sweep.py breaktime,cornell 16 "$s1" ${s2:+"$s2"} ${s3:+"$s3"} >> $out 2>&1
done
cat $out
