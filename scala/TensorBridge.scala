// `Tensor` <-> native memory bridge of the B200 host runtime (reference: core/Tensor.scala:13-32,159-183; core/TensorCodec.scala:12-41;
// native/RawTensors.scala:24-30).  A reference `Tensor` is a read-only view over a direct ByteBuffer that aliases TF-owned memory
// (C-contiguous, native byte order, 4 bytes per Float): the engine reads it through a zero-copy MemorySegment over that very
// buffer (`segment`), and parameters come back through one sb_params_get into a direct buffer and `Tensor.fromBytes`.
// NOT COMPILED in this repository's CI (no JVM in the build image).
package scanet.b200

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.nio.ByteOrder

import scanet.core.{Floating, Params, Path, Shape, Tensor, TensorType}
import scanet.models.{Model, ParamDef}
import scanet.optimizers.{Algorithm, StepResult}

object TensorBridge {

  /** Zero-copy view of a tensor's memory for a downcall.  `toByteBuffer` compacts a sliced view first (core/Tensor.scala:44), so
    * the segment is always the C-contiguous row-major image the C ABI expects; it stays valid as long as the tensor is reachable. */
  def segment[A](t: Tensor[A]): MemorySegment = {
    val buf = t.toByteBuffer.order(ByteOrder.nativeOrder())
    require(buf.isDirect, "scanet tensors live in native memory (RawTensors.allocate)")
    MemorySegment.ofBuffer(buf)
  }

  /** The engine's tensors by the reference's path strings (sb_param_info): (path, shape, role, aggregation). */
  final case class ParamInfo(path: Path, shape: Shape, role: Int, aggregation: Int)

  def paramInfos(engine: MemorySegment): Seq[ParamInfo] = {
    val n = Native.paramCount.invoke(engine).asInstanceOf[Int]
    val arena = Arena.ofConfined()
    try {
      val path = arena.allocate(256)
      val dims = arena.allocate(JAVA_LONG, 4)
      val rank = arena.allocate(JAVA_INT)
      val role = arena.allocate(JAVA_INT)
      val agg = arena.allocate(JAVA_INT)
      val off = arena.allocate(JAVA_LONG)
      (0 until n).map { i =>
        Native.check(Native.paramInfo.invoke(engine, i, path, 256L, dims, rank, role, agg, off).asInstanceOf[Int])
        val r = rank.get(JAVA_INT, 0)
        val shape = Shape((0 until r).map(j => dims.getAtIndex(JAVA_LONG, j.toLong).toInt): _*)
        ParamInfo(Path.parse(path.getString(0)), shape, role.get(JAVA_INT, 0), agg.get(JAVA_INT, 0))
      }
    } finally arena.close()
  }

  /** One arena tensor as a reference `Tensor` (sb_params_get -> direct buffer -> `Tensor.fromBytes`, core/Tensor.scala:176-180). */
  def readParam[A: Floating: TensorType](engine: MemorySegment, info: ParamInfo): Tensor[A] = {
    val arena = Arena.ofConfined()
    try {
      val n = info.shape.power
      val host = arena.allocate(JAVA_FLOAT, math.max(n, 1).toLong)
      Native.check(Native.paramsGet.invoke(engine, arena.allocateFrom(info.path.toString), host, n.toLong).asInstanceOf[Int])
      Tensor.fromBytes[A](host.asSlice(0, 4L * n).toArray(JAVA_BYTE), info.shape)
    } finally arena.close()
  }

  private val RoleWeight = 0
  private val RoleState = 1
  private val RoleOpt = 2

  /** `TrainedModel.params`: every model tensor (trainable weights and layer state), keyed like `model.params(..)`. */
  def readModelParams[A: Floating: TensorType](engine: MemorySegment, model: Model, batchSize: Int): Params[Tensor[A]] = {
    val single = Descriptors.leaves(model).size == 1 // a single layer keeps un-prefixed paths (Conv2DLayerSpec.scala:35)
    Params(paramInfos(engine).filter(_.role != RoleOpt).map { info =>
      val path = if (single) Path(info.path.segments.tail) else info.path
      path -> readParam[A](engine, info)
    }: _*)
  }

  def readOptParams[A: Floating: TensorType](engine: MemorySegment): Params[Tensor[A]] =
    Params(paramInfos(engine).filter(_.role == RoleOpt).map(info => info.path -> readParam[A](engine, info)): _*)

  /** `StepResult` of an epoch (Optimizer.scala:26-35) whose tensors leave the device only when an effect touches them
    * (`RecordLoss` needs the scalar alone; `RecordAccuracy` reads `modelParams`, optimizers/Effect.scala:52-64). */
  def lazyStepResult[A: Floating: TensorType](
      engine: MemorySegment,
      iterations: Int,
      loss: Float,
      model: Model,
      alg: Algorithm,
      batchSize: Int)(implicit num: Numeric[A]): StepResult[A] = {
    val infos = paramInfos(engine)
    def defs(keep: ParamInfo => Boolean): Params[ParamDef] =
      Params(infos.filter(keep).map { i =>
        i.path -> ParamDef(i.shape, aggregation = if (i.aggregation == 1) Some(scanet.models.Aggregation.Avg) else None,
          trainable = i.role == RoleWeight)
      }: _*)
    lazy val modelParams = readModelParams[A](engine, model, batchSize)
    lazy val optParams = readOptParams[A](engine)
    new StepResult[A](
      iterations,
      defs(_.role != RoleOpt),
      new LazyParams(modelParams),
      defs(_.role == RoleOpt),
      new LazyParams(optParams),
      num.fromInt(0).asInstanceOf[A] match { case _ => Floating[A].fromFloat(loss) })
  }

  /** `Params` whose map is materialised on first access: `Params` is a case class over a strict Map (core/Params.scala:19), so the
    * laziness lives in a by-name constructor argument evaluated once. */
  final class LazyParams[A](make: => Params[Tensor[A]]) extends Params[Tensor[A]](Map.empty) {
    private lazy val real = make
    override def unwrap: Map[Path, Tensor[A]] = real.params
    override val params: Map[Path, Tensor[A]] = Map.empty // shadowed: every reader goes through `unwrap` / `apply`
    override def apply(path: Path): Tensor[A] = real(path)
    override def paths: Set[Path] = real.paths
  }
}
