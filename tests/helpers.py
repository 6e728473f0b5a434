"""Test-only CPU helpers: rebuild an organism's painted set from PACKED sub-tiles so the
host packing can be checked without a GPU.  Lives under tests/ on purpose: the product
package has no CPU evaluation path."""
import numpy as np

BLK = 64

def edge_ids_for_dna(pw, dna):
    """Step -> edge id (CSR position, or E+from for a non-edge); -1 where no step is
    executed.  Mirrors the step rule of mappy.py:343-355 (first -1 or idx==L-1 stops)."""
    dna = np.asarray(dna)
    A, L = dna.shape
    out = -np.ones((A, L), dtype=np.int64)
    for a in range(A):
        for i in range(L):
            w = dna[a, i]
            if w == -1 or i == L - 1:
                break
            nxt = dna[a, i + 1]
            if nxt < 0:
                nxt += pw.N
            lo, hi = int(pw.row_ptr[w]), int(pw.row_ptr[w + 1])
            hit = np.nonzero(pw.col[lo:hi] == nxt)[0]
            out[a, i] = lo + hit[0] if len(hit) else pw.E + w
    return out


def painted_blocks(pw, edge_ids):
    """dict block -> [64] u64 union of the sub-tiles of the given edges."""
    acc = {}
    for e in np.unique(edge_ids[edge_ids >= 0]):
        for s in range(int(pw.sub_ptr[e]), int(pw.sub_ptr[e + 1])):
            b = int(pw.sub_block[s])
            pay = int(pw.sub_payload[s])
            off, qmask = pay >> 4, pay & 15
            blk = acc.setdefault(b, np.zeros(BLK, dtype='<u8'))
            k = 0
            for q in range(4):
                if qmask >> q & 1:
                    blk[16 * q:16 * q + 16] |= pw.tile_data[(off + k) * 16:(off + k + 1) * 16]
                    k += 1
    return acc
