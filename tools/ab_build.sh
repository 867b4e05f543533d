# A/B of build-time variants on the GPU box: rebuild the library with extra nvcc flags (PN_NVCC_FLAGS, read by
# protnote_b200/build.py), then run the sustained sweep.   bash tools/ab_build.sh "" "-DSOME_EXPERIMENT"
for flags in "$@"; do
  echo "=== PN_NVCC_FLAGS=$flags"
  PN_NVCC_FLAGS="$flags" python -m protnote_b200.build --force > /dev/null
  bash tools/enc_sweep.sh "promote_k_encoder=64"
done
