"""Minimal BAM / BAI / BGZF readers for the tests (the image has no samtools / pysam).  Test helper only."""
import struct
import zlib

import numpy as np

SEQ_NT16 = "=ACMGRSVTWYHKDBN"
CIGAR_OPS = "MIDNSHP=X"


def bgzf_blocks(path):
    """-> list of (file_offset, uncompressed bytes); checks every member's CRC32 / ISIZE and the EOF marker"""
    raw = open(path, "rb").read()
    out, p = [], 0
    while p < len(raw):
        assert raw[p:p + 4] == b"\x1f\x8b\x08\x04", f"bad gzip member header at {p}"
        xlen = struct.unpack_from("<H", raw, p + 10)[0]
        assert xlen == 6 and raw[p + 12:p + 14] == b"BC" and struct.unpack_from("<H", raw, p + 14)[0] == 2
        bsize = struct.unpack_from("<H", raw, p + 16)[0] + 1
        data = zlib.decompress(raw[p + 18:p + bsize - 8], -15)
        crc, isize = struct.unpack_from("<II", raw, p + bsize - 8)
        assert zlib.crc32(data) == crc and len(data) == isize and isize <= 65536
        out.append((p, data))
        p += bsize
    assert out and out[-1][1] == b"" and len(raw) - out[-1][0] == 28, "missing BGZF EOF block"
    return out


class Bam:
    def __init__(self, path):
        self.blocks = bgzf_blocks(path)
        self.block_index = {off: i for i, (off, _) in enumerate(self.blocks)}
        starts, tot = [], 0
        for _, d in self.blocks:
            starts.append(tot)
            tot += len(d)
        self.starts = starts
        self.data = b"".join(d for _, d in self.blocks)
        d = self.data
        assert d[:4] == b"BAM\x01"
        l_text = struct.unpack_from("<i", d, 4)[0]
        self.text = d[8:8 + l_text].decode()
        p = 8 + l_text
        n_ref = struct.unpack_from("<i", d, p)[0]
        p += 4
        self.refs = []
        for _ in range(n_ref):
            l_name = struct.unpack_from("<i", d, p)[0]
            name = d[p + 4:p + 4 + l_name - 1].decode()
            l_ref = struct.unpack_from("<i", d, p + 4 + l_name)[0]
            self.refs.append((name, l_ref))
            p += 8 + l_name
        self.records = []
        while p < len(d):
            start = p
            bs = struct.unpack_from("<i", d, p)[0]
            (rid, pos, l_name, mapq, bin_, n_cig, flag, l_seq, mrid, mpos, tlen) = struct.unpack_from("<iiBBHHHiiii", d, p + 4)
            q = p + 36
            name = d[q:q + l_name - 1].decode()
            q += l_name
            cigar = list(struct.unpack_from(f"<{n_cig}I", d, q))
            q += 4 * n_cig
            sb = d[q:q + (l_seq + 1) // 2]
            seq = "".join(SEQ_NT16[(sb[i >> 1] >> (4 if i % 2 == 0 else 0)) & 15] for i in range(l_seq))
            q += (l_seq + 1) // 2
            qual = np.frombuffer(d[q:q + l_seq], dtype=np.uint8)
            q += l_seq
            tags = {}
            end = p + 4 + bs
            while q < end:
                tag, typ = d[q:q + 2].decode(), chr(d[q + 2])
                q += 3
                if typ == "Z":
                    e = d.index(b"\0", q)
                    tags[tag] = d[q:e].decode()
                    q = e + 1
                else:
                    fmt = {"c": "<b", "C": "<B", "s": "<h", "S": "<H", "i": "<i", "I": "<I"}[typ]
                    tags[tag] = struct.unpack_from(fmt, d, q)[0]
                    q += struct.calcsize(fmt)
            assert q == end
            self.records.append(dict(rid=rid, pos=pos, mapq=mapq, bin=bin_, flag=flag, cigar=cigar, mrid=mrid, mpos=mpos,
                                     tlen=tlen, name=name, seq=seq, qual=qual, tags=tags, ustart=start, uend=end))
            p = end

    def voffset_to_u(self, v):
        """virtual offset -> offset in the concatenated uncompressed stream"""
        return self.starts[self.block_index[v >> 16]] + (v & 0xffff)


def read_bai(path):
    d = open(path, "rb").read()
    assert d[:4] == b"BAI\x01"
    n_ref = struct.unpack_from("<i", d, 4)[0]
    p = 8
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", d, p)[0]
        p += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", d, p)
            p += 8
            bins[b] = [struct.unpack_from("<QQ", d, p + 16 * k) for k in range(n_chunk)]
            p += 16 * n_chunk
        n_intv = struct.unpack_from("<i", d, p)[0]
        p += 4
        lin = list(struct.unpack_from(f"<{n_intv}Q", d, p))
        p += 8 * n_intv
        refs.append((bins, lin))
    n_no_coor = struct.unpack_from("<Q", d, p)[0] if p + 8 <= len(d) else None
    return refs, n_no_coor


def read_tbi(path):
    """tabix index -> dict(format, col_seq, col_beg, col_end, meta, skip, names, refs=[(bins, linear)], n_no_coor)"""
    d = b"".join(x for _, x in bgzf_blocks(path))
    assert d[:4] == b"TBI\x01"
    n_ref, fmt, col_seq, col_beg, col_end, meta, skip, l_nm = struct.unpack_from("<8i", d, 4)
    p = 36
    names = d[p:p + l_nm].split(b"\0")[:-1]
    assert len(names) == n_ref
    p += l_nm
    refs = []
    for _ in range(n_ref):
        n_bin = struct.unpack_from("<i", d, p)[0]
        p += 4
        bins = {}
        for _ in range(n_bin):
            b, n_chunk = struct.unpack_from("<Ii", d, p)
            p += 8
            bins[b] = [struct.unpack_from("<QQ", d, p + 16 * k) for k in range(n_chunk)]
            p += 16 * n_chunk
        n_intv = struct.unpack_from("<i", d, p)[0]
        p += 4
        lin = list(struct.unpack_from(f"<{n_intv}Q", d, p))
        p += 8 * n_intv
        refs.append((bins, lin))
    n_no_coor = struct.unpack_from("<Q", d, p)[0] if p + 8 <= len(d) else None
    assert p + (8 if n_no_coor is not None else 0) == len(d)
    return dict(format=fmt, col_seq=col_seq, col_beg=col_beg, col_end=col_end, meta=meta, skip=skip,
                names=[n.decode() for n in names], refs=refs, n_no_coor=n_no_coor)


def reg2bins(beg, end):
    """all bins overlapping [beg, end) (SAM spec 5.3)"""
    end -= 1
    out = [0]
    for shift, off in ((26, 1), (23, 9), (20, 73), (17, 585), (14, 4681)):
        out += list(range(off + (beg >> shift), off + (end >> shift) + 1))
    return out


class BgzfText:
    """random access into a BGZF text file by virtual offset"""

    def __init__(self, path):
        self.blocks = bgzf_blocks(path)
        self.index = {off: i for i, (off, _) in enumerate(self.blocks)}

    def lines_between(self, v0, v1):
        i0, o0 = self.index[v0 >> 16], v0 & 0xffff
        i1, o1 = self.index[v1 >> 16], v1 & 0xffff
        if i0 == i1:
            data = self.blocks[i0][1][o0:o1]
        else:
            data = self.blocks[i0][1][o0:] + b"".join(self.blocks[k][1] for k in range(i0 + 1, i1)) + self.blocks[i1][1][:o1]
        return [ln for ln in data.split(b"\n") if ln]


def tabix_query(tbi, text, name, beg, end):
    """data lines of sequence `name` overlapping [beg, end) (0-based), the way tabix finds them: candidate chunks from the
    bins, trimmed by the linear index, then filtered by coordinates"""
    if name not in tbi["names"]:
        return []
    bins, lin = tbi["refs"][tbi["names"].index(name)]
    min_off = lin[min(beg >> 14, len(lin) - 1)] if lin else 0
    out = []
    for b in reg2bins(beg, end):
        for v0, v1 in bins.get(b, []):
            if v1 <= min_off:
                continue
            for ln in text.lines_between(v0, v1):
                f = ln.split(b"\t")
                if f[0].decode() != name:
                    continue
                p = int(f[1]) - 1
                if p < end and p + len(f[3]) > beg:
                    out.append(ln)
    return sorted(set(out), key=lambda ln: int(ln.split(b"\t")[1]))
