"""One short device-resident run for ncu captures: python tools/dev/prof_case.py <case> [reps]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from quick_bench import run

NAT = dict(mode=1, pct_coded=70)
CASES = {
    "ipb_mc": lambda reps: run("1080p420 IPB prediction only", 1920, 1088, 1, reps=reps, seed=3, gop_n=15, gop_m=3, n_gops=4, mode=1, pct_coded=0, pct_intra_in_pb=0),
    "intra_nat": lambda reps: run("1080p420 intra natural x8", 1920, 1088, 1, reps=reps, seed=2, intra_only=1, gop_n=8, n_gops=1, gop_m=1, natural_mean_coefs=6, **NAT),
    "ipb_nat": lambda reps: run("1080p420 IPB natural 2gops", 1920, 1088, 1, reps=reps, seed=3, gop_n=15, gop_m=3, n_gops=2, natural_mean_coefs=5, **NAT),
    "intra420": lambda reps: run("1080p420 intra x8", 1920, 1088, 1, reps=reps, seed=2, intra_only=1, gop_n=8, n_gops=1),
    "ipb420": lambda reps: run("1080p420 IPB 2gops", 1920, 1088, 1, reps=reps, seed=3, gop_n=15, gop_m=3, n_gops=2),
    "ipb444": lambda reps: run("4k444 IPB 1gop", 3840, 2160, 3, reps=reps, seed=4, gop_n=7, gop_m=3, n_gops=1),
}

if __name__ == "__main__":
    CASES[sys.argv[1]](int(sys.argv[2]) if len(sys.argv) > 2 else 2)
