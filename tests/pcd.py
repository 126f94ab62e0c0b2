"""PCD files for the CLI tests (written independently of the product's reader/writer)."""
import numpy as np

_HEADER = ("# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS {fields}\nSIZE {sizes}\nTYPE {types}\n"
           "COUNT {counts}\nWIDTH {n}\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS {n}\nDATA {kind}\n")


def _lzf_literal_stream(raw: bytes) -> bytes:
    """A valid LZF stream made of literal runs only (ctrl byte n-1 < 32 followed by n bytes)."""
    out = bytearray()
    for i in range(0, len(raw), 32):
        chunk = raw[i:i + 32]
        out.append(len(chunk) - 1)
        out += chunk
    return bytes(out)


def write_pcd(path, cloud, kind="ascii", extra_field=False):
    """cloud: [N,>=3] float array.  extra_field adds an `intensity` column the reader has to skip."""
    xyz = np.ascontiguousarray(cloud[:, :3], dtype=np.float32)
    n = len(xyz)
    cols = [xyz[:, 0], xyz[:, 1], xyz[:, 2]]
    names = ["x", "y", "z"]
    if extra_field:
        cols.insert(1, np.arange(n, dtype=np.float32))   # between x and y on purpose
        names.insert(1, "intensity")
    k = len(names)
    hdr = _HEADER.format(fields=" ".join(names), sizes=" ".join(["4"] * k), types=" ".join(["F"] * k),
                         counts=" ".join(["1"] * k), n=n, kind=kind)
    with open(path, "wb") as f:
        f.write(hdr.encode())
        if kind == "ascii":
            for row in zip(*cols):
                f.write((" ".join("%.9g" % v for v in row) + "\n").encode())
        elif kind == "binary":
            f.write(np.stack(cols, axis=1).astype(np.float32).tobytes())
        elif kind == "binary_compressed":
            raw = b"".join(np.ascontiguousarray(c, dtype=np.float32).tobytes() for c in cols)   # field by field
            packed = _lzf_literal_stream(raw)
            f.write(np.array([len(packed), len(raw)], dtype=np.uint32).tobytes())
            f.write(packed)
        else:
            raise ValueError(kind)


def read_pcd_xyz(path):
    """Reads back what the product's writer produces (FIELDS x y z, ascii or binary)."""
    data = open(path, "rb").read()
    head, _, body = data.partition(b"DATA ")
    kind, _, body = body.partition(b"\n")
    n = int([ln for ln in head.decode().splitlines() if ln.startswith("POINTS")][0].split()[1])
    if kind.strip() == b"ascii":
        vals = np.array(body.decode().split(), dtype=np.float64).reshape(n, 3)
        return vals.astype(np.float32)
    return np.frombuffer(body[:12 * n], dtype=np.float32).reshape(n, 3).copy()
