#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/variants_enc.log
for rep in 1 2; do
  python tools/enc_rate.py 64 32 >> gpurun_out/variants_enc.log 2>&1
  for v in variants/grid*.so; do MILLION_B200_LIB=$PWD/$v python tools/enc_rate.py 64 >> gpurun_out/variants_enc.log 2>&1; done
done
sort gpurun_out/variants_enc.log | uniq
