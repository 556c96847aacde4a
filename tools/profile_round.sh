# ncu passes for profiles/ (run on the GPU box via gpurun; numbers printed under ncu are never bench values)
set -x
B="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --no-strong --segments 0"
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv $B > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:sse_stream_kernel -s 3 -c 1 -o gpurun_out/produce -f $B > /dev/null 2>&1
ncu --set full --import-source on --clock-control none --kernel-name regex:sse_decode_kernel -s 3 -c 1 -o gpurun_out/decode -f $B > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -s 18 -c 6 --csv --log-file gpurun_out/traffic.csv $B > /dev/null 2>&1
ls -la gpurun_out/
