"""Test helper: builds HFDL PDUs octet by octet (MPDU downlink / uplink headers, LPDUs, SPDUs with frame check sequences as
mpdu.c:56-119, lpdu.c:129-150, spdu.c:55-62 read them) and mutates them, for the front-parser tests.  The FCS routine is
passed in: the reference's own crc.c where it is built (tests/test_oracle_front_ref.py), else the oracle's pinned one."""
import numpy as np


def oracle_fcs(buf):
    import ctypes as C
    import orclib as O
    L = O.lib()
    L.orc_crc16.restype = C.c_uint16
    a = np.frombuffer(bytes(buf), np.uint8).copy()
    v = L.orc_crc16(a, a.size, 0xFFFF) ^ 0xFFFF
    return bytes([v & 0xFF, v >> 8])


class Forge:
    def __init__(self, rng, fcs=oracle_fcs):
        self.rng, self.fcs = rng, fcs

    def octets(self, n):
        return bytes(self.rng.integers(0, 256, n, dtype=np.uint8))

    def lpdu(self, n, good=True, typ=0x0D):
        """an LPDU of n octets in all (lpdu.c:129-150: type, payload, 2 FCS octets)"""
        if n < 3:
            return self.octets(n)
        body = bytes([typ]) + self.octets(n - 3)
        f = self.fcs(body)
        if not good:
            f = bytes([f[0] ^ (1 << int(self.rng.integers(0, 8))), f[1]])
        return body + f

    def downlink(self, lens, good=None, hdr_good=True, claim=None):
        """downlink MPDU (mpdu.c:56-59,90-99): octet 0 = 1 | 2 | lpdu_cnt << 2, 6 + cnt header octets, FCS, LPDUs"""
        cnt = len(lens)
        hdr = bytearray(self.octets(6 + cnt))
        hdr[0] = (hdr[0] & 0xC0) | 0x03 | (cnt << 2)
        for j, n in enumerate(lens):
            hdr[6 + j] = (claim[j] if claim else n) - 1
        f = self.fcs(hdr)
        if not hdr_good:
            f = bytes([f[0], f[1] ^ 0x40])
        return bytes(hdr) + f + b"".join(self.lpdu(n, good is None or good[j]) for j, n in enumerate(lens))

    def uplink(self, per_ac, good=None):
        """uplink MPDU (mpdu.c:60-75,100-119): octet 0 = 1 | (aircraft_cnt - 1) << 4, then per aircraft: id, cnt << 4, cnt length octets"""
        r = self.rng
        hdr = bytearray([0x01 | ((len(per_ac) - 1) << 4) | (int(r.integers(0, 2)) << 7), int(r.integers(0, 256))])
        body, k = b"", 0
        for lens in per_ac:
            hdr += bytes([int(r.integers(0, 256)), (len(lens) << 4) | int(r.integers(0, 16))]) + bytes(n - 1 for n in lens)
            for n in lens:
                body += self.lpdu(n, good is None or good[k])
                k += 1
        return bytes(hdr) + self.fcs(hdr) + body

    def spdu(self, good=True, extra=0):
        body = bytearray(self.octets(64))
        body[0] &= 0xFE
        f = self.fcs(body)
        if not good:
            f = bytes([f[0] ^ 0x10, f[1]])
        return bytes(body) + f + self.octets(extra)

    def branches(self):
        """one PDU per branch of the front parser"""
        d, u = self.downlink, self.uplink
        return [
            d([]),                                           # no LPDU at all
            d([3]), d([2]), d([1]),                          # shortest good LPDU, too short ones (lpdu.c:137)
            d([20, 30, 40], good=[True, False, True]),
            d([256]),                                        # longest LPDU (length octet 255)
            d([10] * 15),                                    # most LPDUs a downlink header can announce
            d([10, 10], hdr_good=False),
            d([10, 12], claim=[10, 13]),                     # last LPDU runs one octet past the PDU (mpdu.c:152-156)
            d([10, 12], claim=[11, 12]),                     # first length wrong: both LPDUs misparsed, second truncated
            u([[12]]), u([[12, 9], [15]], good=[True, False, True]),
            u([[5]] * 8),                                    # eight aircraft
            u([[], [4], []]),                                # aircraft without LPDUs
            u([[7] * 15, [9] * 15]),
            u([[12, 9], [15]])[:-4],                         # truncated inside the last LPDU
            u([[12, 9], [15]])[:7],                          # truncated inside the header (mpdu.c:66-70)
            d([5, 5])[:9],                                   # header FCS cut off (mpdu.c:79-83)
            self.spdu(), self.spdu(False), self.spdu()[:65], self.spdu(extra=3),
        ]

    def fuzz(self, count):
        """valid PDUs of every kind and random octets, then truncated / bit-flipped / padded"""
        r = self.rng
        out = []
        for i in range(count):
            k = i % 4
            if k == 0:
                lens = [int(r.integers(1, 40)) for _ in range(int(r.integers(0, 16)))]
                p = self.downlink(lens, good=[bool(r.integers(0, 4)) for _ in lens])
            elif k == 1:
                per = [[int(r.integers(1, 30)) for _ in range(int(r.integers(0, 6)))] for _ in range(int(r.integers(1, 9)))]
                p = self.uplink(per, good=[bool(r.integers(0, 4)) for _ in range(sum(len(x) for x in per))])
            elif k == 2:
                p = self.spdu(extra=int(r.integers(0, 4)))
            else:
                p = self.octets(int(r.integers(1, 200)))
            p = bytearray(p)
            m = int(r.integers(0, 6))
            if m == 0 and len(p) > 1:
                p = p[: int(r.integers(1, len(p)))]                                        # truncation anywhere
            elif m == 1:
                p[int(r.integers(0, min(len(p), 12)))] ^= 1 << int(r.integers(0, 8))       # a bit error in the header
            elif m == 2:
                p[int(r.integers(0, len(p)))] ^= 1 << int(r.integers(0, 8))                # a bit error anywhere
            elif m == 3:
                p += self.octets(int(r.integers(1, 30)))                                   # trailing octets (fill of the slot)
            out.append(bytes(p[:945]))
        return out
