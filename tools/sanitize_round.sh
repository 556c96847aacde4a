# compute-sanitizer passes for profiles/ (run on the GPU box via gpurun). A reduced selection of the parity tests: the tools slow
# kernels down 10-100x. Logs go to gpurun_out/; copy them to profiles/<tag>_{memcheck,racecheck,initcheck}.log.
SEL='tests/test_gpu_parity.py -k "(split or templates or fused) and (appendix or fixtures or edges or carry or reset or tricky or C1 or C2-256-4 or C4-512-6 or mixed)"'
for tool in memcheck racecheck; do
  eval timeout 900 compute-sanitizer --tool $tool --error-exitcode 77 --log-file gpurun_out/$tool.log python -m pytest $SEL -m gpu -x -q > gpurun_out/${tool}_pytest.txt 2>&1
  echo "$tool exit $?" >> gpurun_out/${tool}_pytest.txt
  tail -3 gpurun_out/${tool}_pytest.txt; tail -3 gpurun_out/$tool.log
done
