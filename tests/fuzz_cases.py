"""Deterministic pseudo-random `bwa mem` invocations (options, read length, error model, batch size) for the
differential tests: case k is always the same command line."""
import random


def case(k):
    rng = random.Random(7919 * (k + 1))
    pick = lambda p: rng.random() < p
    o = []
    if pick(.3): o += ["-k", str(rng.choice([11, 15, 19, 23, 30]))]
    if pick(.3): o += ["-w", str(rng.choice([5, 20, 50, 100, 200]))]
    if pick(.2): o += ["-d", str(rng.choice([20, 50, 100, 200]))]
    if pick(.2): o += ["-r", str(rng.choice([0.5, 1.0, 1.5, 3.0]))]
    if pick(.2): o += ["-c", str(rng.choice([2, 20, 500, 10000]))]
    if pick(.2): o += ["-D", str(rng.choice([0.1, 0.5, 0.9]))]
    if pick(.2): o += ["-W", str(rng.choice([0, 10, 40]))]
    if pick(.2): o += ["-m", str(rng.choice([0, 5, 50]))]
    if pick(.15): o += ["-y", str(rng.choice([0, 5, 20, 100]))]
    if pick(.25): o += ["-A", str(rng.choice([1, 2, 3]))]
    if pick(.25): o += ["-B", str(rng.choice([2, 4, 6, 9]))]
    if pick(.2): o += ["-O", rng.choice(["6", "4,8", "10", "1,1"])]
    if pick(.2): o += ["-E", rng.choice(["1", "2,1", "3"])]
    if pick(.2): o += ["-L", rng.choice(["5", "0", "3,7", "20"])]
    if pick(.15): o += ["-U", str(rng.choice([0, 9, 17, 40]))]
    if pick(.2): o += ["-T", str(rng.choice([0, 10, 30, 60]))]
    if pick(.15): o += ["-h", rng.choice(["5", "1,50", "0", "20,200"])]
    for f in "aMY5qSPCVj":
        if pick(.08): o += ["-" + f]
    if pick(.1): o += ["-x", rng.choice(["intractg", "pacbio", "ont2d"])]
    if pick(.1): o += ["-G", str(rng.choice([100, 1000, 100000]))]
    if pick(.1): o += ["-N", str(rng.choice([1, 5, 100]))]
    if pick(.1): o += ["-s", str(rng.choice([5, 10, 50]))]
    if pick(.1): o += ["-z", str(rng.choice([0.5, 0.8, 0.95]))]
    ref = rng.choice(["stress", "stress", "two", "c1"])
    paired = pick(.5)
    if paired and pick(.15): o += ["-I", rng.choice(["400,50", "300,30,800,100"])]
    return dict(ref=ref, paired=paired, length=rng.choice([36, 75, 100, 150, 150, 250, 400]), n=40 if paired else 60, seed=rng.randrange(1 << 30),
                err=rng.choice([(0.008, 0.001, 0.001), (0.016, 0.002, 0.002), (0.03, 0.005, 0.005), (0.0, 0.0, 0.0)]),
                chimeric=rng.choice([0.0, 0.05, 0.2]), opts=o, K=str(rng.choice([100000000, 3000, 20000])))


def command(data, k):
    c = case(k)
    fa, fqs = data.reads(c["ref"], tag="fz%d" % k, n=c["n"], length=c["length"], seed=c["seed"], paired=c["paired"], err=c["err"], chimeric=c["chimeric"])
    return c["opts"] + ["-K", c["K"], "-t", "2", fa] + fqs
