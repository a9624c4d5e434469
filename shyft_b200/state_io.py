"""Cell-identified state: the Python face of api/api_state.h (`model.state.extract_state(cids)`, `.apply_state(states, cids)`,
`PTGSKStateWithIdVector.serialize_to_bytes / deserialize_from_bytes / .state_vector`; api/boostpython/expose.h:44-88,
shyft/api/pt_gs_k/__init__.py:56).  Selection and id matching run behind the C ABI (sb2_extract_state / sb2_apply_state; the
state rows are gathered / scattered on the device).

Byte format: the reference writes boost.serialization binary archives (core/core_serialization.h); boost is not available here, so
`serialize_to_bytes` has its own documented little-endian layout -- NOT interchangeable with the reference's blobs:
    magic "SB2S" | u32 version = 1 | i64 n | i64 state_size | n x (cid, x, y, area) i64 | n x state_size f64
"""
import ctypes as C
import struct

import numpy as np

from . import capi

ID_DTYPE = np.dtype([("cid", np.int64), ("x", np.int64), ("y", np.int64), ("area", np.int64)])
_MAGIC = b"SB2S"


def cell_state_id_of(geo_cells):
    """cell_state_id_of (api/api_state.h:58-60): the catchment id and the int-truncated mid point x, y and area"""
    g = np.asarray(geo_cells)
    ids = np.zeros(g.shape[0], dtype=ID_DTYPE)
    ids["cid"] = g["catchment_id"]
    for k in ("x", "y", "area"):
        ids[k] = np.trunc(g[k]).astype(np.int32).astype(np.int64)
    return ids


class StateWithIdVector:
    """vector<cell_state_with_id<state_t>>: .ids (cid, x, y, area) and .states [n][state_size]"""

    def __init__(self, ids, states):
        self.ids = np.ascontiguousarray(ids, dtype=ID_DTYPE)
        st = np.ascontiguousarray(states, dtype=np.float64)
        self.states = st if st.ndim == 2 else st.reshape(self.ids.shape[0], -1)

    def __len__(self):
        return int(self.ids.shape[0])

    @property
    def state_vector(self):
        """extract_state_vector (api/boostpython/expose.h:44-57): the plain states, ids dropped"""
        return self.states.copy()

    def serialize_to_bytes(self):
        return _MAGIC + struct.pack("<Iqq", 1, len(self), self.states.shape[1]) + self.ids.tobytes() + self.states.tobytes()

    @classmethod
    def deserialize_from_bytes(cls, blob):
        blob = bytes(blob)
        if blob[:4] != _MAGIC:
            raise RuntimeError("not a shyft_b200 cell-state blob")
        version, n, k = struct.unpack_from("<Iqq", blob, 4)
        if version != 1:
            raise RuntimeError(f"unknown cell-state blob version {version}")
        off = 4 + struct.calcsize("<Iqq")
        ids = np.frombuffer(blob, dtype=ID_DTYPE, count=n, offset=off)
        states = np.frombuffer(blob, dtype=np.float64, count=n * k, offset=off + n * ID_DTYPE.itemsize).reshape(n, k)
        return cls(ids.copy(), states.copy())


class StateIoHandler:
    """state_io_handler<cell_t> (api/api_state.h:93-142) of one model"""

    def __init__(self, model):
        self._m = model

    def extract_state(self, cids=()):
        m = self._m
        c = np.ascontiguousarray(cids, dtype=np.int64)
        n, k = m.size(), m.state_size
        ids = np.zeros(n, dtype=ID_DTYPE)
        states = np.zeros((n, k))
        n_out = C.c_int64(0)
        m._ck(m._L.sb2_extract_state(m._h, c.ctypes.data_as(capi.c_i64p), C.c_int(c.size), ids.ctypes.data_as(C.c_void_p), capi.dptr(states),
                                     C.byref(n_out)))
        return StateWithIdVector(ids[:n_out.value], states[:n_out.value])

    def apply_state(self, cell_id_state_vector, cids=()):
        """-> indices (into the supplied vector) of the states that matched `cids` but no cell"""
        m = self._m
        v = cell_id_state_vector
        c = np.ascontiguousarray(cids, dtype=np.int64)
        if len(v) and v.states.shape[1] != m.state_size:
            raise RuntimeError("cell state size does not match the model's stack")
        missing = np.zeros(max(len(v), 1), dtype=np.int64)
        n_missing = C.c_int64(0)
        m._ck(m._L.sb2_apply_state(m._h, C.c_int64(len(v)), v.ids.ctypes.data_as(C.c_void_p), capi.dptr(v.states), c.ctypes.data_as(capi.c_i64p),
                                   C.c_int(c.size), missing.ctypes.data_as(capi.c_i64p), C.byref(n_missing)))
        return missing[:n_missing.value].tolist()
