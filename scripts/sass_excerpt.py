"""Counts of the SASS mnemonics that prove what each kernel of libcgs_b200.so is made of (tcgen05 = UTC*MMA,
tcgen05.ld = LDTM, TMA = UTMALDG / UTMASTG, no-return global reductions = REDG, system-scope peer loads / stores and
flag barriers of the exchange kernels = .SYS, shared-memory atomics = ATOMS / REDS ...).
usage: python scripts/sass_excerpt.py > profiles/r2_sass_excerpt.txt"""
import collections, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "pycistopic_b200", "libcgs_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA(?:\.[A-Z0-9_]+)*|UTCBAR(?:\.[A-Z0-9_]+)*|LDTM(?:\.[A-Za-z0-9_]+)*|STTM(?:\.[A-Za-z0-9_]+)*|"
                 r"UTMALDG(?:\.[A-Z0-9_]+)*|UTMASTG(?:\.[A-Z0-9_]+)*|UBLKCP(?:\.[A-Z0-9_]+)*|SYNCS(?:\.[A-Z0-9_]+)*|"
                 r"REDG(?:\.[A-Z0-9_]+)*|RED(?:\.[A-Z0-9_]+)*|ATOMG(?:\.[A-Z0-9_]+)*|ATOMS(?:\.[A-Z0-9_]+)*|REDS(?:\.[A-Z0-9_]+)*|"
                 r"LDG(?:\.[A-Z0-9_]+)*\.SYS|STG(?:\.[A-Z0-9_]+)*\.SYS|LD(?:\.[A-Z0-9_]+)*\.SYS|ST(?:\.[A-Z0-9_]+)*\.SYS|"
                 r"MEMBAR(?:\.[A-Z0-9_]+)*|HMMA(?:\.[A-Z0-9_]+)*|DFMA|DADD|DMUL|MUFU\.[A-Z0-9]+|SHFL\.[A-Z]+|MATCH\.[A-Z]+|VOTE(?:U)?\.[A-Z]+)\b")
print("# cuobjdump -sass pycistopic_b200/libcgs_b200.so (sm_100a): selected mnemonics per kernel")
print("# built from", subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip())
cur, counts, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = re.sub(r"\(.*", "", cur)
        counts[cur] = collections.Counter()
        continue
    if cur is None or "/*" not in line:
        continue
    total[cur] += 1 if re.search(r"^\s+/\*[0-9a-f]{4}\*/", line) else 0
    for mm in pat.findall(line):
        counts[cur][mm] += 1
want = sys.argv[1:] or ["sweep_kernel<4, 4, float, false>", "sweep_kernel<4, 4, float, true>", "sweep_kernel<4, 4, double, false>",
                        "sweep_kernel<1, 1, float, false>", "sweep_det_kernel<4, 4", "exchange_round_kernel", "exchange_sync_kernel",
                        "delta_apply_kernel", "impute_umma_pair_kernel<true, 5>", "impute_umma_pair_kernel<true, 2>",
                        "rank_columns_kernel<int>", "ranksum_rows_kernel<int>", "count_kernel", "recount_kernel", "topic_gram_kernel"]
for name, c in counts.items():
    short = name.replace("cgs::", "").replace("void ", "")
    if not any(w in short for w in want):
        continue
    print(f"\n{short}   [{total[name]} instructions]")
    for k, v in sorted(c.items(), key=lambda kv: (-kv[1], kv[0])):
        print(f"    {v:6d}  {k}")
