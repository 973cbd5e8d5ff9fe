"""Per-kernel SASS mnemonic counts of the built library (CPU only):
cuobjdump -sass quip_for_all_b200/lib/libquipb200.so > /tmp/all.sass && python tools/sass_summary.py /tmp/all.sass > profiles/rNN_sass_summary.txt
UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, UTMALDG = cp.async.bulk.tensor (TMA tensor map), UBLKCP = cp.async.bulk,
SYNCS = mbarrier, IDP.4A = dp4a, HMMA = mma.sync (legacy tensor path), UCGABAR = cluster barrier."""
import collections
import re
import subprocess
import sys

KEYS = ['UTCHMMA', 'LDTM', 'UTCBAR', 'UTMALDG', 'UBLKCP', 'SYNCS', 'UCGABAR', 'IDP.4A', 'HMMA', 'LDSM', 'REDUX', 'SHFL', 'LDGSTS', 'ATOM', 'RED',
        'MEMBAR', 'LDS', 'STS']
WANT = ('e8p_umma_kernel', 'rotblk_pipe', 'rot4096', 'decode_step_kernel', 'ql_gemv_kernel', 'cluster_kernel', 'lm_tail', 'decompress_e8',
        'attn_decode', 'ql_prologue_kernel', 'ql_epilogue_kernel', 'e8p_nearest', 'handoff_')


def main(path):
    print(__doc__.strip().replace("\n", "\n# ").join(["# ", ""]))
    rows = []
    for blk in open(path).read().split("Function :")[1:]:
        name = blk.split('\n')[0].strip()
        ins = re.findall(r'^\s+/\*[0-9a-f]+\*/\s+(.*?);', blk, flags=re.M)
        c = collections.Counter()
        for i in ins:
            op = re.sub(r'^@!?U?P\d+\s+', '', i).split()[0]
            for k in KEYS:
                if op.startswith(k):
                    c[k] += 1
        d = subprocess.run(['c++filt', name], capture_output=True, text=True).stdout.strip() or name
        if any(k in d for k in WANT):
            rows.append((re.sub(r'\(.*\)$', '', d), len(ins), c))
    for d, n, c in sorted(rows):
        print(f"{d}: {n} instr; " + ", ".join(f"{k} {c[k]}" for k in KEYS if c[k]))


if __name__ == "__main__":
    main(sys.argv[1])
