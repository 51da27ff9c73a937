#!/bin/bash
# compute-sanitizer memcheck over the round's new kernels at tiny dims (non-causal U-Net, windowed streaming flow, fade kernel)
mkdir -p gpurun_out
timeout -k 10 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_unet_nc_gpu.py::test_unet_nc_matches_reference_fixture" "tests/test_unet_nc_gpu.py::test_unet_nc_lengths_vs_oracle" -m gpu -q --timeout 800 -k "tiny or lengths" > gpurun_out/san_unet_nc.log 2>&1
tail -6 gpurun_out/san_unet_nc.log
timeout -k 10 900 compute-sanitizer --tool memcheck --print-limit 20 python -m pytest "tests/test_flow_gpu.py::test_flow_stream_incremental_matches_recompute" tests/test_streaming_gpu.py -m gpu -q --timeout 800 -k "tiny or streaming or cv2 or abandoned" > gpurun_out/san_stream.log 2>&1
tail -6 gpurun_out/san_stream.log
grep -c "Invalid\|out of bounds\|misaligned" gpurun_out/san_unet_nc.log gpurun_out/san_stream.log
