# A/B of build-time variants on the GPU box: rebuild with the given nvcc flags, then run the sustained sweep
for flags in "$@"; do
  echo "=== PN_NVCC_FLAGS=$flags"
  PN_NVCC_FLAGS="$flags" python -m protnote_b200.build --force > /dev/null
  bash tools/enc_sweep.sh "promote_k_encoder=64"
done
