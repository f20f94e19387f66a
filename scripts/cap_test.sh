#!/bin/bash
for cap in 16 24 48; do
  if [ $cap != 48 ]; then cp yunikorn_k8shim_b200/libykgpu.so /tmp/orig.so; cp gpurun_cap$cap.so yunikorn_k8shim_b200/libykgpu.so; fi
  echo CAP $cap; YK_PROFILE_COMMIT=1 python scripts/quick_time.py 2>&1 | tail -6 | grep -o "batch=[0-9]*\|commit=[0-9.]*ms\|prof=.*" | paste - - - | sed -n '2p;5p'
  if [ $cap != 48 ]; then cp /tmp/orig.so yunikorn_k8shim_b200/libykgpu.so; fi
done
