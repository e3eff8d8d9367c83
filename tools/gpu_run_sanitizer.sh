#!/bin/bash
# compute-sanitizer over the product kernels: memcheck / racecheck / synccheck / initcheck on smoke() and on a
# subset of the parity tests that covers every kernel and epilogue (ragged + padded layout, int16 + float PCM,
# utterance / global CMVN, masks before / after, features-in path, dither, 48 k ingest, one-call batch entry).
cd "${GRAFT_REPO_ROOT:-.}"
SEL='test_fixture_fbank_vs_reference_golden or test_fixture_float32_pcm_path or test_fixture_utterance_cmvn or test_cmvn_arithmetic_in_isolation or test_specaugment_class_vs_reference_golden or test_fused_fbank_cmvn_specaugment_batch or test_edge_lengths or test_global_cmvn_two_pass or test_tile_boundary_utterances_in_padded_layout or test_float_pcm_two_slot_path or test_dither_compat or test_batch_call_matches or test_reformat'
for tool in memcheck racecheck synccheck initcheck; do
  echo "==== compute-sanitizer --tool $tool: smoke()"
  timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -4
  echo "exit ${PIPESTATUS[0]}"
  echo "==== compute-sanitizer --tool $tool: pytest subset"
  timeout 1500 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider 2>&1 | tail -6
  echo "exit ${PIPESTATUS[0]}"
done
