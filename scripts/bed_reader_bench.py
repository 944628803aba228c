"""Reader throughput (host only): cgs_bed_read against the two engines of the reference's read_fragments_to_pyranges
that exist in this image (pandas C parser, pyarrow), on a synthetic position-sorted fragments file.
usage: python scripts/bed_reader_bench.py [n_fragments] [n_barcodes] [dir]"""
import gzip
import os
import struct
import sys
import time
import zlib

import numpy as np
import pandas as pd

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pycistopic_b200 import fragment_matrix as fm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
n_bc = int(sys.argv[2]) if len(sys.argv) > 2 else 20_000
out = sys.argv[3] if len(sys.argv) > 3 else "/tmp"
rng = np.random.default_rng(0)
chroms = np.array([f"chr{i}" for i in range(1, 23)] + ["chrX"])
bc = np.array(["".join(x) + "-1" for x in rng.choice(list("ACGT"), (n_bc, 16))])
c = np.sort(rng.integers(0, len(chroms), n))
start = rng.integers(0, 150_000_000, n)
order = np.lexsort((start, c))
c, start = c[order], start[order]
end = start + rng.integers(30, 800, n)
df = pd.DataFrame({"c": chroms[c], "s": start, "e": end, "b": bc[rng.integers(0, n_bc, n)], "n": rng.integers(1, 4, n)})
plain, gz = os.path.join(out, "bench_fragments.tsv"), os.path.join(out, "bench_fragments.tsv.gz")
df.to_csv(plain, sep="\t", header=False, index=False)
bgz = os.path.join(out, "bench_fragments.bgzf.tsv.gz")
with open(plain, "rb") as f, gzip.open(gz, "wb", compresslevel=3) as g, open(bgz, "wb") as h:
    while chunk := f.read(1 << 24):
        g.write(chunk)
        for i in range(0, len(chunk), 65280):  # BGZF members, as bgzip writes them
            part = chunk[i:i + 65280]
            c = zlib.compressobj(3, zlib.DEFLATED, -15)
            body = c.compress(part) + c.flush()
            h.write(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(body) + 25)
                    + body + struct.pack("<II", zlib.crc32(part) & 0xFFFFFFFF, len(part)))
    h.write(bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000"))
size = os.path.getsize(plain)
print(f"{n} fragments, {n_bc} barcodes, {size / 1e6:.0f} MB plain, {os.path.getsize(gz) / 1e6:.0f} MB gzip, "
      f"{os.path.getsize(bgz) / 1e6:.0f} MB bgzip, {os.cpu_count()} host cores; best of two runs each")


def timed(label, fn, bytes_):
    dt = 1e9
    for _ in range(2):  # the first touch of fresh memory is slow on a VM: whoever runs first would pay for the others
        t = time.perf_counter()
        r = fn()
        dt = min(dt, time.perf_counter() - t)
    print(f"{label:48s} {dt:7.2f} s  {n / dt / 1e6:7.2f} M fragments/s  {bytes_ / dt / 1e6:7.0f} MB/s (plain text)")
    return r


names = ["Chromosome", "Start", "End", "Name", "Score"]
dtypes = {"Chromosome": "category", "Start": np.int32, "End": np.int32, "Name": "category"}
for path in (plain, gz, bgz):
    tag = "bgzip" if path == bgz else "gzip " if path == gz else "plain"
    t = timed(f"{tag} cgs_bed_read (all cores) -> DataFrame", lambda: fm.read_fragments_from_file(path), size)
    timed(f"{tag} cgs_bed_read (1 thread) -> DataFrame", lambda: fm.read_fragments_from_file(path, n_threads=1), size)
    r = timed(f"{tag} reference engine 'pandas' (read_table)", lambda: pd.read_table(
        path, sep="\t", header=None, names=names, doublequote=False, engine="c", dtype=dtypes), size)
    assert (t["Start"].to_numpy() == r["Start"].to_numpy()).all() and list(t["Name"].iloc[:1000].astype(str)) == list(
        r["Name"].iloc[:1000].astype(str))
    import pyarrow as pa
    import pyarrow.csv
    timed(f"{tag} reference engine 'pyarrow' (threads)", lambda: pa.csv.read_csv(
        path, read_options=pa.csv.ReadOptions(use_threads=True, column_names=names),
        parse_options=pa.csv.ParseOptions(delimiter="\t", quote_char=False),
        convert_options=pa.csv.ConvertOptions(column_types={
            "Chromosome": pa.dictionary(pa.int32(), pa.string()), "Start": pa.int32(), "End": pa.int32(),
            "Name": pa.dictionary(pa.int32(), pa.string())})).to_pandas(), size)
os.remove(plain)
os.remove(gz)
os.remove(bgz)

# ---- duplicates: a four-column file (no Score), a third of the records repeated --------------------------------------
# the reference collapses them per chromosome with utils.collapse_duplicates (np.lexsort over an object array,
# cistopic_class.py:819-829); here: cgs_bed_collapse_duplicates inside the reader
n4 = min(n, 3_000_000)
dup = df.iloc[:n4, :4]
dup = pd.concat([dup, dup.iloc[rng.integers(0, n4, n4 // 3)]], ignore_index=True)
four = os.path.join(out, "bench_fragments4.tsv")
dup.to_csv(four, sep="\t", header=False, index=False)
size4, n_all = os.path.getsize(four), len(dup)
n = n_all
print(f"{n_all} records without Score, {n4 // 3} of them duplicates, {size4 / 1e6:.0f} MB plain")
got = timed("plain cgs_bed_read + cgs_bed_collapse_duplicates", lambda: fm.read_bed_table(four, 4, collapse_duplicates=True), size4)


def reference_way():
    import pyarrow as pa
    import pyarrow.csv
    f = pa.csv.read_csv(four, read_options=pa.csv.ReadOptions(use_threads=True, column_names=names[:4]),
                        parse_options=pa.csv.ParseOptions(delimiter="\t", quote_char=False),
                        convert_options=pa.csv.ConvertOptions(column_types={
                            "Chromosome": pa.dictionary(pa.int32(), pa.string()), "Start": pa.int32(), "End": pa.int32(),
                            "Name": pa.dictionary(pa.int32(), pa.string())})).to_pandas()

    def collapse(d):  # utils.py:284-293, verbatim arithmetic
        a = d.values
        sidx = np.lexsort(a[:, :4].T)
        b = a[sidx, :4]
        m = np.concatenate(([True], (b[1:] != b[:-1]).any(1), [True]))
        return pd.DataFrame(np.column_stack((b[m[:-1], :4], np.diff(np.flatnonzero(m) + 1))),
                            columns=["Chromosome", "Start", "End", "Name", "Score"])
    return pd.concat([collapse(f[f.Chromosome == x]) for x in f.Chromosome.cat.categories.values])


ref = timed("plain reference: pyarrow engine + collapse_duplicates", reference_way, size4)
assert len(ref) == len(got["start"]) and int(ref["Score"].sum()) == int(got["score"].sum()) == n_all
os.remove(four)
