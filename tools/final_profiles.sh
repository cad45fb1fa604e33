set -x
timeout 700 python -m pytest tests -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.log
timeout 300 python bench.py --steps 100 --warmup 10 > gpurun_out/r02_bench_1gpu_final.json 2> gpurun_out/r02_bench_1gpu_final.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_ncu_launches_final.csv python bench.py --steps 2 --warmup 3 > gpurun_out/r02_ncu_launches_final.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_kernel_bn_fwd -s 2 -c 1 -o gpurun_out/r02_ncu_bn_fwd_stage1 -f python tools/fused_layer_once.py > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:wgrad_gemm_kernel -s 2 -c 1 -o gpurun_out/r02_ncu_wgrad_stage1b -f python tools/fused_layer_once.py > /dev/null 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:conv_gemm_kernel_bn_bwd -s 31 -c 1 -o gpurun_out/r02_ncu_bn_bwd_stage1 -f python tools/bwd_layer_once.py > /dev/null 2>&1
timeout 200 python tools/layer_timeline.py > gpurun_out/r02_layer_timelines_final.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
