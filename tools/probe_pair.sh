#!/bin/bash
# dev helper: tcgen05 GEMM probe, CTA pairs on vs off, guarded against hangs
for pm in 1 0; do
  B200_PAIR=$pm timeout 180 python tools/probe_tc.py fwd dgrad > gpurun_out/probe_pair$pm.log 2>&1; echo "rc=$?" >> gpurun_out/probe_pair$pm.log
  echo "== pair mode $pm"; cut -c1-175 gpurun_out/probe_pair$pm.log | sed 's/mma *[0-9.]*us |//g'
done
