#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 120 python tests/debug_patch_embed.py > gpurun_out/dbg_pe.log 2>&1; echo "exit $?" >> gpurun_out/dbg_pe.log
DBG_PDL=1 timeout 120 python tests/debug_patch_embed.py > gpurun_out/dbg_pe_pdl.log 2>&1; echo "exit $?" >> gpurun_out/dbg_pe_pdl.log
timeout 300 compute-sanitizer --tool memcheck python tests/debug_patch_embed.py > gpurun_out/dbg_pe_memcheck.log 2>&1; echo "exit $?" >> gpurun_out/dbg_pe_memcheck.log
tail -8 gpurun_out/dbg_pe.log; tail -25 gpurun_out/dbg_pe_memcheck.log | cut -c1-200
