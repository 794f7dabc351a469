// Panama (java.lang.foreign, JDK 22+) binding of include/scanet_b200.h -- the JVM side of the drop-in boundary.
// NOT COMPILED in this repository's CI: the build image has no JVM (DESIGN.md section 1).  Every downcall below mirrors one
// prototype of include/scanet_b200.h; struct layouts follow the header field by field.
package scanet.b200

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.lang.invoke.MethodHandle

object Native {
  private val linker = Linker.nativeLinker()
  private val arena = Arena.global()
  private val lib = SymbolLookup.libraryLookup(
    sys.props.getOrElse("scanet.b200.lib", "libscanet_b200.so"), arena)

  private def fn(name: String, ret: MemoryLayout, args: MemoryLayout*): MethodHandle =
    linker.downcallHandle(
      lib.find(name).orElseThrow(() => new UnsatisfiedLinkError(s"$name not exported by libscanet_b200.so")),
      FunctionDescriptor.of(ret, args: _*))

  // ---- sb_initializer / sb_layer_desc / sb_model_desc / sb_train_desc (include/scanet_b200.h:78-141) ----
  val Initializer: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("kind"), JAVA_INT.withName("has_seed"), JAVA_LONG.withName("seed"),
    JAVA_FLOAT.withName("a"), JAVA_FLOAT.withName("b"))
  val LayerDesc: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("kind"), JAVA_INT.withName("units"), JAVA_INT.withName("activation"),
    JAVA_INT.withName("recurrent_activation"), JAVA_FLOAT.withName("relu_threshold"),
    JAVA_INT.withName("window_h"), JAVA_INT.withName("window_w"), JAVA_INT.withName("stride_h"),
    JAVA_INT.withName("stride_w"), JAVA_INT.withName("padding"), JAVA_INT.withName("pool_reduce"),
    JAVA_INT.withName("use_bias"), JAVA_INT.withName("trainable"), JAVA_FLOAT.withName("bn_momentum"),
    JAVA_FLOAT.withName("bn_epsilon"), JAVA_INT.withName("bn_mode"), JAVA_INT.withName("bn_fused"),
    MemoryLayout.paddingLayout(4), MemoryLayout.sequenceLayout(4, Initializer).withName("init"),
    JAVA_INT.withName("reg_kind"), JAVA_FLOAT.withName("reg_lambda"),
    JAVA_INT.withName("pad_top"), JAVA_INT.withName("pad_bottom"), JAVA_INT.withName("pad_left"), JAVA_INT.withName("pad_right"),
    JAVA_INT.withName("return_sequence"), JAVA_INT.withName("pool_honor_reduce"), JAVA_INT.withName("stateful"))
  val ModelDesc: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("n_layers"), MemoryLayout.paddingLayout(4), ADDRESS.withName("layers"),
    JAVA_INT.withName("input_rank"), MemoryLayout.paddingLayout(4),
    MemoryLayout.sequenceLayout(4, JAVA_LONG).withName("input_dims"),
    JAVA_INT.withName("label_rank"), MemoryLayout.paddingLayout(4),
    MemoryLayout.sequenceLayout(4, JAVA_LONG).withName("label_dims"))
  val TrainDesc: StructLayout = MemoryLayout.structLayout(
    JAVA_INT.withName("batch"), JAVA_INT.withName("loss"), JAVA_INT.withName("optimizer"),
    JAVA_FLOAT.withName("rate"), JAVA_FLOAT.withName("beta1"), JAVA_FLOAT.withName("beta2"),
    JAVA_FLOAT.withName("epsilon"), JAVA_FLOAT.withName("init_acc"), JAVA_INT.withName("nesterov"),
    JAVA_INT.withName("minimizing"), JAVA_INT.withName("precision"), JAVA_INT.withName("use_graph"))

  // ---- entry points ----
  val lastError: MethodHandle = fn("sb_last_error", ADDRESS)
  val deviceCount: MethodHandle = fn("sb_device_count", JAVA_INT)
  val engineCreate: MethodHandle = fn("sb_engine_create", JAVA_INT, ADDRESS, ADDRESS, JAVA_INT, ADDRESS)
  val engineDestroy: MethodHandle =
    linker.downcallHandle(lib.find("sb_engine_destroy").get(), FunctionDescriptor.ofVoid(ADDRESS))
  val paramCount: MethodHandle = fn("sb_param_count", JAVA_INT, ADDRESS)
  val paramInfo: MethodHandle =
    fn("sb_param_info", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val paramsSet: MethodHandle = fn("sb_params_set", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG)
  val paramsGet: MethodHandle = fn("sb_params_get", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG)
  val initParams: MethodHandle = fn("sb_init_params", JAVA_INT, ADDRESS)
  val initOptParams: MethodHandle = fn("sb_init_opt_params", JAVA_INT, ADDRESS)
  val datasetUpload: MethodHandle = fn("sb_dataset_upload", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG)
  val trainEpoch: MethodHandle = fn("sb_train_epoch", JAVA_INT, ADDRESS, JAVA_LONG, ADDRESS, ADDRESS)
  val trainStep: MethodHandle = fn("sb_train_step", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
  val trainBatchesAsync: MethodHandle =
    fn("sb_train_batches_async", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG, ADDRESS)
  val evalLoss: MethodHandle = fn("sb_eval_loss", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val predict: MethodHandle = fn("sb_predict", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG)
  // estimators/package.scala:18-137 on the device: (engine, kind, x|NULL, y|NULL, rows, double[4])
  val estimate: MethodHandle = fn("sb_estimate", JAVA_INT, ADDRESS, JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
  // KryoSerializers.scala:27-63 records <-> arena tensors, and whole-arena checkpoints
  val paramToWire: MethodHandle = fn("sb_param_to_wire", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
  val paramFromWire: MethodHandle = fn("sb_param_from_wire", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, JAVA_LONG, ADDRESS)
  val checkpointSave: MethodHandle = fn("sb_checkpoint_save", JAVA_INT, ADDRESS, ADDRESS, JAVA_LONG, JAVA_LONG)
  val checkpointLoad: MethodHandle = fn("sb_checkpoint_load", JAVA_INT, ADDRESS, ADDRESS, ADDRESS, ADDRESS)
  val groupCreate: MethodHandle = fn("sb_group_create", JAVA_INT, ADDRESS, JAVA_INT, JAVA_INT, ADDRESS, JAVA_INT, ADDRESS)
  val groupAverage: MethodHandle = fn("sb_group_average", JAVA_INT, ADDRESS, ADDRESS, ADDRESS)
  val groupDestroy: MethodHandle =
    linker.downcallHandle(lib.find("sb_group_destroy").get(), FunctionDescriptor.ofVoid(ADDRESS))

  /** sb_status -> the exception the reference throws at the same place (core/Require.scala:5-6, native/package.scala:6-9). */
  def check(status: Int): Unit = if (status != 0) {
    val msg = lastError.invoke().asInstanceOf[MemorySegment].reinterpret(4096).getString(0)
    status match {
      case 1 => throw new IllegalArgumentException(msg)
      case 2 => throw new UnsupportedOperationException(msg)
      case _ => throw new RuntimeException(msg)
    }
  }
}
