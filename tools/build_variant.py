# coding: utf-8
"""Build a tuning variant of the CUDA library next to the product build:
     python tools/build_variant.py NAME -DJS2T_FOO=1 ...   ->  build/libjs2t_NAME.so
   and run anything against it with  JS2T_LIB=build/libjs2t_NAME.so python bench.py ..."""
import subprocess
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from joeys2t_b200 import _lib  # noqa: E402

name, flags = sys.argv[1], sys.argv[2:]
out = Path(__file__).resolve().parent.parent / "build" / f"libjs2t_{name}.so"
out.parent.mkdir(exist_ok=True)
cmd = _lib.nvcc_command(out=out, extra=tuple(flags) + ("-Xptxas", "-v"))
res = subprocess.run(cmd, capture_output=True, text=True)
if res.returncode != 0:
    sys.exit(res.stdout + res.stderr)
for line in res.stderr.splitlines():
    if "fbank_tile_kernel" in line or ("Used" in line and "registers" in line and "barriers" in line):
        print(line)
print(out)
