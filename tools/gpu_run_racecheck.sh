cd "${GRAFT_REPO_ROOT:-.}"
SEL='test_fixture_fbank_vs_reference_golden or test_fixture_float32_pcm_path or test_fixture_utterance_cmvn or test_cmvn_arithmetic_in_isolation or test_specaugment_class_vs_reference_golden or test_fused_fbank_cmvn_specaugment_batch or test_edge_lengths or test_global_cmvn_two_pass or test_tile_boundary_utterances_in_padded_layout or test_float_pcm_two_slot_path or test_dither_compat or test_batch_call_matches or test_reformat or fused_global'
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests -m gpu -x -q -k "$SEL" -p no:cacheprovider > gpurun_out/r2_racecheck_full.txt 2>&1
echo "exit $?"
grep -c "Race reported" gpurun_out/r2_racecheck_full.txt
grep "Race reported\|and .* access" gpurun_out/r2_racecheck_full.txt | sort | uniq -c | sort -rn | head -20
tail -4 gpurun_out/r2_racecheck_full.txt
