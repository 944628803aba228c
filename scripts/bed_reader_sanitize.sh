#!/bin/bash
# The host-side BED / fragments reader (pycistopic_b200/csrc/bed_reader.cpp) under AddressSanitizer + UBSan and under
# ThreadSanitizer: plain, gzip, BGZF, a BGZF file cut inside a member, an empty file and a malformed record, read with
# 1 / 3 / 8 threads and text blocks of 16 B, 5 kB and 32 MB (CGS_BED_BLOCK_BYTES), duplicates collapsed, everything exported.
# usage: scripts/bed_reader_sanitize.sh [workdir]   (writes the report to stdout)
set -e
cd "$(dirname "$0")/.."
W=${1:-/tmp/bed_sanitize}
mkdir -p "$W"
cat > "$W/stub.cpp" <<'EOS'
#include <string>
namespace cgs { thread_local std::string g_last_error; }
extern "C" const char *cgs_last_error(void) { return cgs::g_last_error.c_str(); }
EOS
python - "$W" <<'EOS'
import gzip, struct, sys, zlib
import numpy as np
w = sys.argv[1]
def bgzf(raw, block):
    out = []
    for i in range(0, len(raw), block):
        part = raw[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(part) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(body) + 25)
                   + body + struct.pack("<II", zlib.crc32(part) & 0xFFFFFFFF, len(part)))
    return b"".join(out) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
rng = np.random.default_rng(1)
lines = [f"chr{rng.integers(1, 5)}\t{rng.integers(0, 1000)}\t{rng.integers(1000, 1100)}\tBC{rng.integers(0, 50)}" for _ in range(30000)]
raw = ("#hdr\n\n" + "\n".join(lines)).encode()
open(f"{w}/s4.tsv", "wb").write(raw)
open(f"{w}/s4.bgz", "wb").write(bgzf(raw, 3000))
open(f"{w}/s4.gz", "wb").write(gzip.compress(raw))
open(f"{w}/s4cut.bgz", "wb").write(bgzf(raw, 3000)[:50000])
open(f"{w}/empty.tsv", "wb").write(b"")
open(f"{w}/bad.tsv", "wb").write(b"chr1\t1\t2\tA\nchr1\tx\t2\tA\n")
EOS
for san in address,undefined thread; do
  g++ -g -O1 -fsanitize=$san -std=c++17 -Wno-unknown-pragmas -I/usr/local/cuda/include -x c++ \
      pycistopic_b200/csrc/bed_reader.cpp "$W/stub.cpp" scripts/bed_reader_driver.cpp -o "$W/drv_$san" -lz -lpthread
  for blk in 16 5000 33554432; do
    echo "== -fsanitize=$san CGS_BED_BLOCK_BYTES=$blk"
    CGS_BED_BLOCK_BYTES=$blk "$W/drv_$san" "$W"/s4.tsv "$W"/s4.bgz "$W"/s4.gz "$W"/s4cut.bgz "$W"/empty.tsv "$W"/bad.tsv 2>&1
  done
done
echo "== done (any sanitizer report would be printed above)"
