"""Reader/writer for the "aiodump" container written by oracle/ref_dump.cpp: repeated records
``<name> <count>\\n`` followed by <count> little-endian float64 (test infrastructure)."""
import numpy as np


def read_dump(path):
    out = {}
    with open(path, "rb") as f:
        while True:
            line = f.readline()
            if not line:
                break
            name, n = line.decode().rsplit(" ", 1)
            n = int(n)
            out[name] = np.frombuffer(f.read(8 * n), dtype="<f8").copy()
            assert out[name].size == n, f"short record {name} in {path}"
    return out


def write_dump(path, records):
    with open(path, "wb") as f:
        for name, v in records.items():
            v = np.ascontiguousarray(v, dtype="<f8").ravel()
            f.write(f"{name} {v.size}\n".encode())
            f.write(v.tobytes())


def scalar(d, key):
    return float(d[key][0])
