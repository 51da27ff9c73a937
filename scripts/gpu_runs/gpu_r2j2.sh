#!/bin/bash
mkdir -p gpurun_out
timeout -k 10 600 python scripts/first_audio.py 8 2>&1 | grep "^\[lib" | tee gpurun_out/r2j2_first_audio.log
HVX_LIB_PATH=$PWD/flowmirror_hydravox_b200/libhydravox_b200_oldgemm.so timeout -k 10 600 python scripts/first_audio.py 8 2>&1 | grep "^\[lib" | tee -a gpurun_out/r2j2_first_audio.log
