"""tcgen05 / TMEM / TMA mnemonic counts per kernel of the built library (profiles/rNN_sass_conv_tc.txt).
    python tools/sass_counts.py [path/to/libeegldm.so] > profiles/r02b_sass_conv_tc.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "synthetic-sleep-eeg-signal-generation-using-latent-diffusion-models_b200", "eegldm", "libeegldm.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
names = {}
for line in subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout.splitlines():
    pass
PAT = re.compile(r"\b(UTCHMMA(?:\.2CTA)?|UTCQMMA|LDTM(?:\.x\d+)?|STTM|UBLKCP|UTMALDG(?:\.\dD)?(?:\.2CTA)?|UTMASTG|UTCBAR(?:\.2CTA)?(?:\.MULTICAST)?|SYNCS|USETMAXREG|UTCATOMSWS|ELECT|LDG\.E\.[A-Z0-9.]*256[A-Z.]*)\b")
cur, counts, order = None, {}, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        order.append(cur)
        continue
    if cur:
        for t in PAT.findall(line):
            t = re.sub(r"^LDG\.E\..*256.*$", "LDG.256", t)
            counts[cur][t] += 1
dem = subprocess.run(["cu++filt"] + order, capture_output=True, text=True).stdout.splitlines() if order else []
print("# cuobjdump -sass libeegldm.so (sm_100a): tcgen05 / TMEM / TMA mnemonics per kernel (tools/sass_counts.py)")
print("# UTCHMMA = tcgen05.mma kind::f16 (.2CTA = cta_group::2), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk, UTMALDG = cp.async.bulk.tensor,")
print("# UTCBAR = tcgen05.commit, SYNCS = mbarrier ops, USETMAXREG = setmaxnreg, LDG.256 = 256-bit global loads")
for fn, d in zip(order, dem if len(dem) == len(order) else order):
    c = counts[fn]
    if not any(k.startswith(("UTC", "LDTM", "UBLKCP", "UTMA")) for k in c):
        continue
    d = re.sub(r"eegldm::\(anonymous namespace\)::|eegldm::<unnamed>::", "", d)
    print(f"{d[:110]:110s} " + " ".join(f"{k}={v}" for k, v in sorted(c.items())))
