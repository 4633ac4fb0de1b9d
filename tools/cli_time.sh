#!/bin/bash
# Wall-clock of the command line on a 1 GiB tmpfs file (one process per call).
python - <<'PY'
import sys, time, subprocess; sys.path.insert(0,'.')
from lightweight_snappy_b200 import corpus, api
import torch
corpus.make_corpus("mixed", 1<<30, device="cuda").cpu().numpy().tofile("/dev/shm/in.bin")
del torch
for args in (["-c", "/dev/shm/in.bin", "/dev/shm/out.snp"],) * 2 + (["-d", "/dev/shm/out.snp", "/dev/shm/back.bin"],) * 3 + (["-d", "-r", "/dev/shm/out.snp", "/dev/shm/back.bin"],):
    t = time.perf_counter(); subprocess.run([api.CLI_PATH] + list(args), check=True); print(args[0], f"{time.perf_counter() - t:.3f} s", flush=True)
PY
cmp /dev/shm/in.bin /dev/shm/back.bin && echo same
rm -f /dev/shm/in.bin /dev/shm/out.snp /dev/shm/back.bin
