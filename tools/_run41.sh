python -m pytest tests/test_gpu_trie.py -m gpu -x -q 2>&1 | tail -2
for lib in libgt_pad4.so libgenlm_trie_b200.so; do
  echo "=== lib=$lib"
  GT_LIB_NAME=$lib python tools/chain_times.py 2>&1 | sed -n 2,9p
  GT_LIB_NAME=$lib python tools/trace_tile.py 2>&1 | grep -A11 "later items" | grep "ELL\|item total\|E emit\|E wait"
done
