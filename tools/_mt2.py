import sys
sys.argv=[sys.argv[0]]
exec(open("tools/mode_time.py").read().split('run("raw (mode 0')[0])
run("raw (mode 0, no stats)", lambda p: (p.set_masks(None), p.set_cmvn("none")))
run("global CMVN epilogue (mode 2), no masks", lambda p: (p.set_masks(None), p.set_cmvn("global"), p.set_global_stats(mean, istd)))
run("global CMVN epilogue (mode 2), const-fill masks", lambda p: (p.set_cmvn("global"), p.set_global_stats(mean, istd), masks(p, 0.0)))
