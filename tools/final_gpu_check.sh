#!/bin/bash
# One bounded GPU call at the end of a round (about 5 minutes of box time): the GPU test suite first, then ncu
# evidence for the kernels that have no round-2 capture yet, then (if time is left) a short default bench line.
# Everything is written under gpurun_out/ so that a cut-off call still leaves what it finished.
#   gpurun --timeout 330 -- 'bash tools/final_gpu_check.sh'
mkdir -p gpurun_out
T0=$SECONDS
left() { echo $(( ${BUDGET:-320} - (SECONDS - T0) )); }
echo "[check] pytest -m gpu"
timeout 200 python -m pytest tests -m gpu -q > gpurun_out/r2_final_gputest.log 2>&1
echo "[check] pytest rc=$? after $((SECONDS - T0)) s"; tail -3 gpurun_out/r2_final_gputest.log
if [ "$(left)" -gt 90 ]; then
  echo "[check] plain ncu_targets 2000"
  timeout 60 python tools/ncu_targets.py 2000 > gpurun_out/r2_plain_targets_2000.log 2>&1 && tail -1 gpurun_out/r2_plain_targets_2000.log &&
  timeout $(( $(left) - 20 )) ncu --profile-from-start off --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k regex:"update_q_melt_kernel|update_b_kernel|amg_spgemm_table_kernel|spmv_sell_kernel<.int.[0-2], float, .int.1>" -c 16 -f \
    -o gpurun_out/r2_nodal_spgemm_spmvf python tools/ncu_targets.py 2000 > gpurun_out/r2_ncu_targets.log 2>&1
  echo "[check] ncu rc=$? after $((SECONDS - T0)) s"
fi
if [ "$(left)" -gt 75 ]; then
  echo "[check] short bench"
  timeout $(( $(left) - 5 )) python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2_final_bench.json 2> gpurun_out/r2_final_bench.err
  echo "[check] bench rc=$? after $((SECONDS - T0)) s"; cut -c1-400 gpurun_out/r2_final_bench.json
fi
echo "[check] done after $((SECONDS - T0)) s"
