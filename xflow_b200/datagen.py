"""Deterministic synthetic data in the reference's input format.

Text rows are what LoadData::load_minibatch_hash_data_fread parses
(src/io/load_data_from_disk.cc:103-210):  "<label>\\t<field>:<id>:1 <field>:<id>:1 ...\\n"
with ids printed in decimal; the feature key the model sees is std::hash of the id string.

The generator is a counter-based splitmix64 stream, so the same (seed, shape) gives the same
bytes on every machine and numpy version (fixtures do not need to be committed).
"""
import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def splitmix64(x):
    """Vectorised splitmix64 finaliser over uint64 arrays."""
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        return x ^ (x >> np.uint64(31))


def uniform_u64(seed, n, stream=0):
    with np.errstate(over="ignore"):
        ctr = np.arange(n, dtype=np.uint64) + np.uint64(seed) * np.uint64(0x632BE59BD9B4E019) + \
            np.uint64(stream) * np.uint64(0xD1342543DE82EF95)
    return splitmix64(splitmix64(ctr))


def make_ids(seed, rows, nnz_per_row, id_space, dist="uniform", zipf_s=1.05, ragged=False):
    """Returns (row_ptr uint32[rows+1], ids uint64[nnz], labels uint8[rows])."""
    if ragged:
        lens = (uniform_u64(seed, rows, stream=3) % np.uint64(2 * nnz_per_row)).astype(np.int64)
    else:
        lens = np.full(rows, nnz_per_row, np.int64)
    row_ptr = np.zeros(rows + 1, np.uint32)
    row_ptr[1:] = np.cumsum(lens).astype(np.uint32)
    nnz = int(row_ptr[-1])
    u = uniform_u64(seed, nnz, stream=1)
    if dist == "uniform":
        ids = u % np.uint64(id_space)
    elif dist == "zipf":
        # inverse-CDF of a continuous power law on [1, id_space], exponent zipf_s (> 1)
        x = (u >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))
        a = 1.0 - zipf_s
        ids_f = ((id_space ** a - 1.0) * x + 1.0) ** (1.0 / a)
        ids = np.minimum(ids_f.astype(np.uint64), np.uint64(id_space - 1))
    else:
        raise ValueError(dist)
    labels = ((uniform_u64(seed, rows, stream=2) % np.uint64(4)) == 0).astype(np.uint8)  # Bernoulli(0.25)
    return row_ptr, ids.astype(np.uint64), labels


def write_text(path, row_ptr, ids, labels):
    """Write the reference's 3-field text format; field = token index within the row."""
    with open(path, "w", newline="") as f:
        for r in range(labels.size):
            s, e = int(row_ptr[r]), int(row_ptr[r + 1])
            toks = " ".join("%d:%d:1" % (j, int(ids[s + j])) for j in range(e - s))
            f.write("%d\t%s\n" % (int(labels[r]), toks))


def make_csr_keys(seed, rows, nnz_per_row, id_space, hash_fn, dist="uniform", zipf_s=1.05, ragged=False):
    """CSR batch whose keys are hash_fn(decimal ids) — what the loader would emit for write_text()."""
    row_ptr, ids, labels = make_ids(seed, rows, nnz_per_row, id_space, dist, zipf_s, ragged)
    return row_ptr, hash_fn(ids), labels
