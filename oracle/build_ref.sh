#!/usr/bin/env bash
# ORACLE — TEST / BASELINE INFRASTRUCTURE ONLY.
#
# Installs the UNMODIFIED reference package (joeynmt, from /root/reference) into oracle/_ref so that
# it travels to the GPU box with the repo snapshot (oracle/_ref is git-ignored, not gpurun-ignored).
# Run in the build container only (the GPU box has no /root/reference and uses the installed copy):
#
#     bash oracle/build_ref.sh
#
# What uses it (never the product path — joeys2t_b200/ must not import it):
#   * bench.py --impl reference and the cpu_baseline leg: the reference's own
#     extract_fbank_features -> CMVN, one process per host core       (oracle/ref_runner.py)
#   * tests/test_gpu_reference_callers.py: joeys2t_b200.install() patched into the reference's own
#     SpeechDataset / make_iter / SpeechStreamDataset, compared with tests/golden/ref_batches.npz
# The test fixtures of the reference (ten LibriSpeech excerpts + TSV + vocab files) are copied next to
# the package as data, because its dataset classes read them from disk.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${JS2T_REFERENCE_ROOT:-/root/reference}"
DST="$HERE/_ref"
if [ ! -f "$REF/joeynmt/helpers_for_audio.py" ]; then
  echo "build_ref.sh: $REF is not mounted; keeping whatever is in $DST" >&2
  exit 0
fi
TMP="$(mktemp -d /tmp/js2t_ref_src.XXXXXX)"
trap 'rm -rf "$TMP"' EXIT
# the source tree is read-only and setuptools writes build/ and *.egg-info into it: install from a copy
cp -r "$REF/." "$TMP/src"
rm -rf "$DST"
mkdir -p "$DST"
python -m pip install --quiet --no-index --no-build-isolation --no-deps \
  --find-links /opt/wheelhouse --target "$DST" "$TMP/src"
# fixtures the reference's own dataset classes read (test/data/speech: wavs, TSVs, vocab)
mkdir -p "$DST/test_data"
cp -r "$REF/test/data/speech" "$DST/test_data/speech"
python - "$DST" <<'PY'
import hashlib, pathlib, sys
dst = pathlib.Path(sys.argv[1])
ref = pathlib.Path("/root/reference/joeynmt")
# the installed files must be byte-identical to the reference sources ("unmodified")
bad = [p.name for p in ref.glob("*.py")
       if hashlib.sha256(p.read_bytes()).digest() != hashlib.sha256((dst / "joeynmt" / p.name).read_bytes()).digest()]
assert not bad, bad
print(f"oracle/_ref: joeynmt installed unmodified ({len(list(ref.glob('*.py')))} modules verified byte-identical)")
PY
