import sys
sys.argv=[sys.argv[0]]
exec(open("tools/mode_time.py").read().split('run("raw (mode 0')[0])
run("stats only (mode 0 + finalize)", lambda p: (p.set_masks(None), p.set_cmvn("stats")))
run("utterance CMVN in-kernel finalize (mode 1)", lambda p: (p.set_masks(None), p.set_cmvn("utterance"), p.set_option("fused_cmvn", 1)))
