#!/usr/bin/env bash
# ncu source-level capture of the device Huffman kernel (decode workload, 1,024 images).
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_jpeg_huffman -s 1 -c 1 -f -o /tmp/huff python bench.py --workload decode --steps 1 --warmup 1 > gpurun_out/huff_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/huff_ncu.log
cp /tmp/huff.ncu-rep gpurun_out/huff.ncu-rep; ls -la gpurun_out/huff.ncu-rep
