"""lightweight-snappy_b200: a B200-native (sm_100a) Snappy block codec behind the C API of
tturturiello/lightweight-snappy.  The codec itself is `csrc/` (CUDA kernels + C-ABI,
built into libsnappy_b200.so); `api` is the thin ctypes host mirror used by tests and
bench.py; `corpus` generates the synthetic workloads."""
