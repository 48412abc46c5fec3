"""SASS evidence per kernel of libmmumap_b200.so (run here, no GPU needed):
    python scripts/sass_counts.py > profiles/r02_sass_counts.txt
Counts the mnemonics that prove the Blackwell paths: UTCHMMA (tcgen05.mma), UTMALDG (TMA tensor loads), LDTM (tcgen05.ld),
UTCBAR (tcgen05.commit), SYNCS (mbarrier), REDG (red.global), LDGSTS (cp.async), multimem (ld_reduce / st)."""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multimodal-umap_b200", "libmmumap_b200.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA.2CTA", "UTCHMMA", "UTMALDG", "LDTM", "UTCBAR", "UTCATOMSWS", "SYNCS", "REDG.E.ADD.F32x4", "REDG.E.ADD.F32x2", "REDG",
        "LDGSTS", "MULTIMEM", "LD.E.MULTIMEM", "ST.E.MULTIMEM", "MUFU.EX2", "MUFU.LG2", "ATOMG"]
cur, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[cur]["_total"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)}: instruction counts per kernel (only kernels with at least one listed mnemonic)")
print("# a prefix count includes its longer forms (UTCHMMA includes UTCHMMA.2CTA; REDG includes the .F32x4 / .F32x2 vector forms)")
for name, c in counts.items():
    hits = {k: v for k, v in c.items() if k != "_total"}
    if hits:
        print(f"{name}\n    total {c['_total']}: " + ", ".join(f"{k} x{v}" for k, v in sorted(hits.items())))
