#!/bin/bash
# compute-sanitizer (memcheck, then racecheck) over a reduced GPU test subset: every NTT pass structure
# (shared-memory exchanges), the Merkle kernels (tree_top_kernel's cross-level reads), the partial-sponge
# hashing, the in-place column slots and the quotient interpreter with its cp.async loads.
# Logs land in gpurun_out/; tools/sanitizer_summary.py turns them into profiles/*.md.
cd "${GRAFT_REPO_ROOT:-$(dirname "$0")/..}"
SUBSET='fft_matches_oracle or ifft_matches_oracle or merkle_tree_new_parity or leaf_hash_domain or pipelined_upload_with_partial or extend_columns or from_values_cols or fri_committed_trees_parity or batch_merkle_tree_parity or poseidon_kat'
PLONK='quotient_polys_match_oracle or partial_products_and_zs or full_proof_bytes_match_oracle'
for tool in memcheck racecheck; do
  echo "== $tool: tests/test_gpu_parity.py -k \"$SUBSET\""
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --log-file gpurun_out/sanitizer_${tool}_parity.log \
      python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SUBSET" 2>&1 | tail -3
  echo "rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_parity.log
  echo "== $tool: tests/test_gpu_plonk.py -k \"$PLONK\""
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 86 --log-file gpurun_out/sanitizer_${tool}_plonk.log \
      python -m pytest tests/test_gpu_plonk.py -m gpu -x -q -k "$PLONK" 2>&1 | tail -3
  echo "rc=$?"; tail -4 gpurun_out/sanitizer_${tool}_plonk.log
done
