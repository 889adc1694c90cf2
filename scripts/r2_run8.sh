#!/bin/bash
mkdir -p gpurun_out
OODB200_LIB=ood_in_object_detection_b200/variants/prof.so python scripts/pipe_prof.py > gpurun_out/r2_pipe_prof.json 2> gpurun_out/r2_pipe_prof.err; cat gpurun_out/r2_pipe_prof.json; tail -3 gpurun_out/r2_pipe_prof.err
