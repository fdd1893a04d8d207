"""Star catalogue handling on the host side.

Mirror of the input half of src/StarMap.hs:

* ``read_ppm``        -- readMap, src/StarMap.hs:45-58 (PPM binary catalogue records)
* ``star_color``      -- starColor, src/StarMap.hs:60-72
* ``ra_dec_to_cartesian`` -- src/StarMap.hs:74-75

plus ``synthetic_catalogue``: neither the PPM catalogue nor a stars.kdt is shipped with
the reference (README.md:23 tells the user to download it), and this image has no
network, so benchmarks and tests use a deterministic synthetic catalogue written in the
exact PPM record layout (N = 468 861, seed 20190412).  Anyone with
GHC can feed the same bytes to the real ``generate-tree`` + ``blackstar``.
"""
from __future__ import annotations

import numpy as np

#: numpy view of ``bsb_star`` (include/blackstar_b200.h), 48 bytes
STAR_DTYPE = np.dtype([("pos", "<f8", (3,)), ("hue", "<f8"), ("sat", "<f8"),
                       ("mag", "<i4"), ("pad_", "<i4")], align=False)
assert STAR_DTYPE.itemsize == 48

PPM_HEADER_BYTES = 28
PPM_RECORD_BYTES = 28

DEFAULT_N_STARS = 468_861
DEFAULT_SEED = 20190412

_SPECTRAL = {  # src/StarMap.hs:64-72
    ord("O"): (0.631, 0.39), ord("B"): (0.628, 0.33), ord("A"): (0.622, 0.21),
    ord("F"): (0.650, 0.03), ord("G"): (0.089, 0.09), ord("K"): (0.094, 0.29),
    ord("M"): (0.094, 0.56),
}


def star_color(ch: int):
    """starColor (src/StarMap.hs:63-72): spectral letter -> (hue, saturation)."""
    return _SPECTRAL.get(int(ch), (0.0, 0.0))


def ra_dec_to_cartesian(ra, dec):
    """src/StarMap.hs:74-75"""
    ra = np.asarray(ra, dtype=np.float64)
    dec = np.asarray(dec, dtype=np.float64)
    return np.stack([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)], axis=-1)


_PPM_REC = np.dtype([("ra", ">f8"), ("dec", ">f8"), ("spectral", "u1"), ("pad0", "u1"),
                     ("mag", ">i2"), ("pad1", "u1", (8,))])
assert _PPM_REC.itemsize == PPM_RECORD_BYTES


def read_ppm(data: bytes) -> np.ndarray:
    """readMap (src/StarMap.hs:45-58) followed by starColor' (:60-61), as
    readTreeFromFile applies it (:85).  Returns an array of STAR_DTYPE."""
    if len(data) < PPM_HEADER_BYTES:
        raise ValueError("too few bytes")  # cereal: "too few bytes" from `skip 28`
    n = (len(data) - PPM_HEADER_BYTES) // PPM_RECORD_BYTES
    rec = np.frombuffer(data, dtype=_PPM_REC, count=n, offset=PPM_HEADER_BYTES)
    out = np.zeros(n, dtype=STAR_DTYPE)
    out["pos"] = ra_dec_to_cartesian(rec["ra"].astype("<f8"), rec["dec"].astype("<f8"))
    out["mag"] = rec["mag"].astype(np.int32)
    hue = np.zeros(256)
    sat = np.zeros(256)
    for ch, (h, s) in _SPECTRAL.items():
        hue[ch], sat[ch] = h, s
    out["hue"] = hue[rec["spectral"]]
    out["sat"] = sat[rec["spectral"]]
    return out


def _splitmix64(seed: int, idx: np.ndarray) -> np.ndarray:
    """Element ``idx`` of the splitmix64 stream started at ``seed`` (vectorised, wrapping)."""
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + (idx.astype(np.uint64) + np.uint64(1)) * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return z


def _uniform(seed: int, idx: np.ndarray) -> np.ndarray:
    return (_splitmix64(seed, idx) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def synthetic_catalogue(n: int = DEFAULT_N_STARS, seed: int = DEFAULT_SEED) -> bytes:
    """Deterministic synthetic catalogue in PPM binary layout.

    Directions uniform on the sphere (z ~ U(-1,1), phi ~ U(0,2pi) -> Dec = asin z, RA = phi);
    magnitude m drawn from a density proportional to 10^(0.5 m) on [-1.5, 13.0] and stored
    as round(100 m) (int16); spectral letter from {O,B,A,F,G,K,M,other} with weights
    {.01,.10,.20,.20,.20,.20,.08,.01}.
    """
    i = np.arange(n, dtype=np.uint64)
    u_z = _uniform(seed, 4 * i + 0)
    u_phi = _uniform(seed, 4 * i + 1)
    u_m = _uniform(seed, 4 * i + 2)
    u_s = _uniform(seed, 4 * i + 3)
    dec = np.arcsin(2.0 * u_z - 1.0)
    ra = 2.0 * np.pi * u_phi
    lo, hi = 10.0 ** (0.5 * -1.5), 10.0 ** (0.5 * 13.0)
    m = 2.0 * np.log10(lo + u_m * (hi - lo))
    mag = np.clip(np.rint(100.0 * m), -150, 1300).astype(np.int16)
    letters = np.frombuffer(b"OBAFGKM?", dtype=np.uint8)
    edges = np.cumsum([0.01, 0.10, 0.20, 0.20, 0.20, 0.20, 0.08])
    spectral = letters[np.searchsorted(edges, u_s, side="right")]
    rec = np.zeros(n, dtype=_PPM_REC)
    rec["ra"], rec["dec"], rec["spectral"], rec["mag"] = ra, dec, spectral, mag
    header = b"blackstar-b200 synthetic PPM".ljust(PPM_HEADER_BYTES, b"\0")[:PPM_HEADER_BYTES]
    return header + rec.tobytes()


def synthetic_stars(n: int = DEFAULT_N_STARS, seed: int = DEFAULT_SEED) -> np.ndarray:
    """``read_ppm(synthetic_catalogue(n, seed))``"""
    return read_ppm(synthetic_catalogue(n, seed))


# ---------------------------------------------------------------------------------------- stars.kdt
def _put_char(c: int) -> bytes:
    return chr(c).encode("utf-8")          # cereal's Char encoding is UTF-8


def write_kdt(pos: np.ndarray, mag: np.ndarray, spectral: np.ndarray) -> bytes:
    """What ``generate-tree`` writes (src/StarMap.hs:87-91): ``encode (build toList stars)``.

    The tree is built as kdt-0.2.4 builds it (sort by the axis of the level, axes cycling x, y, z; the element
    at index n ``div`` 2 becomes the node) and encoded as cereal's Generic instances do -- both RECALLED, not
    verified (no GHC here; see csrc/host_setup.cpp: parse_kdt).  Used by the tests and by anyone who wants a
    tree file from a catalogue without running the Haskell tool."""
    import struct
    import sys
    out = bytearray(b"\x00\x00")            # the two function fields (src/StarMap.hs:33-40)
    pos = np.asarray(pos, dtype=np.float64)
    sys.setrecursionlimit(max(10000, sys.getrecursionlimit()))

    def node(idx: np.ndarray, depth: int):
        if len(idx) == 0:
            out.append(1)                    # Empty
            return
        ax = depth % 3
        order = idx[np.argsort(pos[idx, ax], kind="stable")]
        m = len(order) // 2
        k = int(order[m])
        out.append(0)                        # TreeNode
        node(order[:m], depth + 1)
        out.extend(struct.pack(">3d", *pos[k]))
        out.extend(struct.pack(">q", int(mag[k])))
        out.extend(_put_char(int(spectral[k])))
        out.extend(struct.pack(">d", float(pos[k, ax])))
        node(order[m + 1:], depth + 1)

    node(np.arange(len(pos)), 0)
    out.extend(struct.pack(">q", len(pos)))
    return bytes(out)


def catalogue_to_kdt(ppm: bytes) -> bytes:
    """``generate-tree PPM stars.kdt`` (app/GenerateTree.hs:11-29) in one call."""
    n = (len(ppm) - PPM_HEADER_BYTES) // PPM_RECORD_BYTES
    rec = np.frombuffer(ppm, dtype=_PPM_REC, count=n, offset=PPM_HEADER_BYTES)
    pos = ra_dec_to_cartesian(rec["ra"].astype("<f8"), rec["dec"].astype("<f8"))
    return write_kdt(pos, rec["mag"].astype(np.int64), rec["spectral"])
