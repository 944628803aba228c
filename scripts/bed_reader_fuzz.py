"""Mutation fuzzing of the BED / fragments reader under ASan + UBSan: 360 damaged plain / gzip / BGZF files (bytes
changed, cut out, inserted; files truncated) x three block sizes x 1, 3, 8 threads through the sanitizer build of
scripts/bed_reader_sanitize.sh (run that first).  Every file must end in a table or in an error code -- never in a crash
or a sanitizer report.   usage: scripts/bed_reader_sanitize.sh && python scripts/bed_reader_fuzz.py"""
import gzip
import os
import random
import struct
import subprocess
import zlib


def bgzf_compress(raw, block=65280):
    out = []
    for i in range(0, len(raw), block):
        part = raw[i:i + block]
        c = zlib.compressobj(6, zlib.DEFLATED, -15)
        body = c.compress(part) + c.flush()
        out.append(b"\x1f\x8b\x08\x04\0\0\0\0\0\xff" + struct.pack("<H", 6) + b"BC" + struct.pack("<HH", 2, len(body) + 25)
                   + body + struct.pack("<II", zlib.crc32(part) & 0xFFFFFFFF, len(part)))
    return b"".join(out) + bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


random.seed(7)
base = ("#h\n" + "".join(f"chr{i%3}\t{i*10}\t{i*10+5}\tBC{i%7}\t{i%4}\n" for i in range(400))).encode()
variants = {"plain": base, "gz": gzip.compress(base), "bgzf": bgzf_compress(base, block=700)}
d = "/tmp/fuzzdir"; os.makedirs(d, exist_ok=True)
paths = []
for name, raw in variants.items():
    for k in range(120):
        b = bytearray(raw)
        for _ in range(random.randint(1, 6)):
            op = random.random()
            pos = random.randrange(len(b))
            if op < 0.4: b[pos] = random.randrange(256)
            elif op < 0.6: del b[pos:pos + random.randint(1, 50)]
            elif op < 0.8: b[pos:pos] = bytes(random.randrange(256) for _ in range(random.randint(1, 20)))
            else: b = b[:pos]
            if not b: b = bytearray(b"\n")
        p = f"{d}/{name}{k}"
        open(p, "wb").write(bytes(b)); paths.append(p)
bad = 0
for blk in ("16", "300", "33554432"):
    env = dict(os.environ, CGS_BED_BLOCK_BYTES=blk)
    for i in range(0, len(paths), 40):
        r = subprocess.run(["/tmp/bed_sanitize/drv_address,undefined"] + paths[i:i+40], env=env, capture_output=True, text=True)
        if r.returncode != 0 or "Sanitizer" in r.stderr or "runtime error" in r.stderr:
            bad += 1; print("PROBLEM", blk, i, r.returncode, r.stderr[-800:])
print("fuzz batches with problems:", bad)
