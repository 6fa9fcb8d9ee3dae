"""Reader for TensorFlow "tensor bundle" checkpoints (the `variables/variables.index` +
`variables.data-00000-of-00001` pair inside a SavedModel), without TensorFlow.

The reference ships its two trained policies as Keras SavedModels
(example_policies/example_no_obs/models/model-00030, example_with_obs/models/model-00032; loader:
main/testing/test_dqn.py:14-29, `tf.keras.models.load_model`).  TensorFlow is not part of this
stack, so the weights of the 5 -> 64 -> 64 -> 3 network (main/impl/dqn.py:301-314) are read straight
from the bundle: the index is a LevelDB-format table (prefix-compressed keys in blocks, a footer
with the index block's handle) whose values are BundleEntryProto messages {dtype, shape, offset,
size} pointing into the data file.

`load_q_network(path)` returns the network as a torch module: with `policies.GreedyQPolicy` and
`policies.evaluate` it replays a shipped policy on the batched GPU env (artifact row f-4 of
SURVEY.md section 8f).  tools/make_policy_fixtures.py stores the two shipped weight sets as small
.npz fixtures (tests/golden/policy_*.npz) for the GPU box, where the reference tree does not exist.
"""
import os
import struct

import numpy as np

MAGIC = 0xdb4775248b80fb57
DTYPES = {1: np.float32, 2: np.float64, 3: np.int32, 9: np.int64}  # tensorflow DataType enum


def _varint(buf, pos):
    out, shift = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        out |= (b & 0x7f) << shift
        if not b & 0x80:
            return out, pos
        shift += 7


def _block(buf, offset, size):
    """Entries (key, value) of one table block."""
    data = buf[offset:offset + size]
    if buf[offset + size] != 0:
        raise ValueError("compressed table blocks are not supported")
    n_restarts = struct.unpack("<I", data[-4:])[0]
    end = len(data) - 4 - 4 * n_restarts
    pos, key, out = 0, b"", []
    while pos < end:
        shared, pos = _varint(data, pos)
        non_shared, pos = _varint(data, pos)
        vlen, pos = _varint(data, pos)
        key = key[:shared] + data[pos:pos + non_shared]
        pos += non_shared
        out.append((key, data[pos:pos + vlen]))
        pos += vlen
    return out


def _proto(buf):
    """Minimal protobuf wire decoder -> {field: [values]} (varint, fixed32/64, bytes)."""
    pos, out = 0, {}
    while pos < len(buf):
        tag, pos = _varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v = struct.unpack("<Q", buf[pos:pos + 8])[0]
            pos += 8
        elif wt == 2:
            n, pos = _varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack("<I", buf[pos:pos + 4])[0]
            pos += 4
        else:
            raise ValueError("wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def read_bundle(prefix):
    """`prefix` = .../variables/variables -> {key: ndarray} for the dense tensors of the bundle."""
    with open(prefix + ".index", "rb") as f:
        idx = f.read()
    with open(prefix + ".data-00000-of-00001", "rb") as f:
        data = f.read()
    footer = idx[-48:]
    if struct.unpack("<Q", footer[-8:])[0] != MAGIC:
        raise ValueError("not a LevelDB-format table")
    pos = 0
    _, pos = _varint(footer, pos)   # metaindex handle
    _, pos = _varint(footer, pos)
    ioff, pos = _varint(footer, pos)
    isize, pos = _varint(footer, pos)
    tensors = {}
    for _, handle in _block(idx, ioff, isize):
        boff, p = _varint(handle, 0)
        bsize, p = _varint(handle, p)
        for key, val in _block(idx, boff, bsize):
            if not key:
                continue  # BundleHeaderProto
            e = _proto(val)
            dtype = e.get(1, [0])[0]
            if dtype not in DTYPES or 7 in e:  # strings (the object graph), sliced tensors: not needed
                continue
            shape = []
            if 2 in e:
                for dim in _proto(e[2][0]).get(2, []):
                    shape.append(_proto(dim).get(1, [0])[0])
            off, size = e.get(4, [0])[0], e.get(5, [0])[0]
            arr = np.frombuffer(data[off:off + size], dtype=DTYPES[dtype]).reshape(shape)
            tensors[key.decode()] = arr.copy()
    return tensors


def dense_weights(tensors):
    """Kernels [in, out] and biases of the three Dense layers (dqn.py:308-313), in order."""
    out = []
    for layer in range(3):
        base = "layer_with_weights-%d/" % layer
        out.append((tensors[base + "kernel/.ATTRIBUTES/VARIABLE_VALUE"], tensors[base + "bias/.ATTRIBUTES/VARIABLE_VALUE"]))
    return out


def load_weights(path):
    """`path`: a SavedModel directory (contains variables/variables.index), or an .npz written by
    tools/make_policy_fixtures.py -> [(kernel [in, out], bias)] x 3."""
    if os.path.isdir(path):
        return dense_weights(read_bundle(os.path.join(path, "variables", "variables")))
    z = np.load(path)
    return [(z["w%d" % k], z["b%d" % k]) for k in range(3)]


def load_q_network(path, device=None):
    """The reference's Q-network (dqn.py:301-314: Dense 64 relu, Dense 64 relu, Dense 3) with the
    weights of a shipped policy, as a torch module taking normalised observations [N, 5]
    (main/impl/utils.py:23-27; `AquaVecEnv.observe(normalized=True)`)."""
    import torch
    from .dqn import create_model
    weights = load_weights(path)
    net = create_model(weights[0][0].shape[0], weights[2][0].shape[1])
    linears = [m for m in net if isinstance(m, torch.nn.Linear)]
    with torch.no_grad():
        for lin, (w, b) in zip(linears, weights):
            lin.weight.copy_(torch.as_tensor(np.ascontiguousarray(w.T)))  # Keras kernel is [in, out]
            lin.bias.copy_(torch.as_tensor(b))
    return net.to(device) if device is not None else net
