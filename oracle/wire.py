"""TEST INFRASTRUCTURE ONLY -- restatement of the reference's Kryo `TensorSerializer` byte layout
(optimizers/KryoSerializers.scala:27-63) with Python's `struct`.

Kryo 4.0.2 (Spark 3.3.1 ships chill 0.10.0 -> Kryo 4.0.2; `build.sbt:16`): `Output.writeInt(int)` = 4 bytes big-endian,
`Output.writeLongs(long[])` = 8 bytes big-endian each, `Output.writeBytes` = raw.  `tensor.toBytes` is the direct buffer as
is (core/Tensor.scala:46-52): row-major, native (little) endian.  The reference's own spec only round-trips through Kryo
(optimizers/KryoSerializersSpec.scala), it holds no byte-level golden vector: "parity unpinned" at the byte level."""
import struct

import numpy as np

DT_FLOAT = 1  # org.tensorflow.proto.framework.DataType.DT_FLOAT.getNumber (core/TensorType.scala:14,171)


def tensor_write(t: np.ndarray) -> bytes:
    t = np.array(t, dtype="<f4", order="C")  # (ascontiguousarray would promote a scalar to rank 1)
    payload = t.tobytes()
    return (struct.pack(">ii", DT_FLOAT, t.ndim) + struct.pack(f">{t.ndim}q", *t.shape) + struct.pack(">i", len(payload))
            + payload)


def tensor_read(buf: bytes):
    code, rank = struct.unpack_from(">ii", buf, 0)
    assert code == DT_FLOAT
    dims = struct.unpack_from(f">{rank}q", buf, 8)
    (n,) = struct.unpack_from(">i", buf, 8 + 8 * rank)
    off = 12 + 8 * rank
    return np.frombuffer(buf[off:off + n], dtype="<f4").reshape(dims).copy(), off + n
