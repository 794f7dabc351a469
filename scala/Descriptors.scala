// Lowers the reference's layer DSL to the flat C descriptors of include/scanet_b200.h (sb_model_desc / sb_layer_desc /
// sb_train_desc): the JVM-side twin of scanet3_b200/dsl.py `Layer.to_native` and scanet3_b200/engine.py `Engine.__init__`.
// A model is walked the way `Composed` walks it (layer/Composed.scala:89-99: nested compositions are flattened, leaves keep their
// position as the integer path prefix `0/`, `1/`, ... -- layer/Composed.scala:18-27), so leaf i of the descriptor list is the owner of
// the parameter paths "i/...".
// NOT COMPILED in this repository's CI (no JVM in the build image); kept self-consistent with B200Optimizer.scala / Native.scala.
package scanet.b200

import java.lang.foreign._
import java.lang.foreign.ValueLayout._

import scanet.core.{Floating, Shape, Tensor}
import scanet.math.nn.ConvFormat.NHWC
import scanet.math.nn.Padding
import scanet.math.nn.Reduce
import scanet.models.Activation._
import scanet.models.Initializer._
import scanet.models.Loss._
import scanet.models.Regularization.{L1, L2, Zero}
import scanet.models.layer._
import scanet.models.{Activation, Initializer, Loss, Model, Regularization}
import scanet.optimizers._

object Descriptors {

  // enums of include/scanet_b200.h
  object Kind { val Dense = 1; val Bias = 2; val Activate = 3; val Flatten = 4; val Conv2D = 5; val Pool2D = 6; val BatchNorm = 7; val LSTM = 8; val SimpleRNN = 9 }
  object Act { val Identity = 0; val Sigmoid = 1; val Tanh = 2; val Softmax = 3; val ReLU = 4 }
  object Pad { val Valid = 0; val Same = 1; val Explicit = 2 }
  object Init {
    val Zeros = 0; val Ones = 1; val Const = 2; val RandomUniform = 3; val RandomNormal = 4; val GlorotUniform = 5; val HeUniform = 6
    val LecunUniform = 7; val Orthogonal = 8; val TruncatedNormal = 9; val GlorotNormal = 10; val HeNormal = 11; val LecunNormal = 12
  }

  /** One leaf of the flattened model: fields of `sb_layer_desc` that differ from the defaults. */
  final case class Leaf(
      kind: Int,
      units: Int = 0,
      activation: Int = Act.Identity,
      recurrentActivation: Int = Act.Sigmoid,
      reluThreshold: Float = 0f,
      window: (Int, Int) = (1, 1),
      strides: (Int, Int) = (1, 1),
      padding: Int = Pad.Valid,
      explicitPad: (Int, Int, Int, Int) = (0, 0, 0, 0), // top, bottom, left, right
      poolReduce: Int = 0,
      useBias: Boolean = false,
      trainable: Boolean = true,
      bnMomentum: Float = 0.99f,
      bnEpsilon: Float = 1e-3f,
      bnFused: Int = -1,
      init: Seq[Initializer] = Seq.empty, // slots 0..3 of sb_layer_desc.init
      reg: Regularization = Zero,
      returnSequence: Boolean = false,
      stateful: Boolean = false)

  private def act(a: Activation): (Int, Float) = a match {
    case Identity   => (Act.Identity, 0f)
    case Sigmoid    => (Act.Sigmoid, 0f)
    case Tanh       => (Act.Tanh, 0f)
    case Softmax    => (Act.Softmax, 0f)
    case ReLU(t)    => (Act.ReLU, t)
    case other      => throw new UnsupportedOperationException(s"activation $other is outside the accelerated path")
  }

  /** The `Composed` leaf walk (layer/Composed.scala:89-99): nested compositions flatten, everything else is one leaf. */
  def leaves(model: Model): Seq[Layer] = model match {
    case Composed(layers) => layers.flatMap(leaves)
    case layer: Layer     => Seq(layer)
    case other            => throw new UnsupportedOperationException(s"$other is not a layer model")
  }

  def lower(layer: Layer): Leaf = layer match {
    case d: Dense => Leaf(Kind.Dense, units = d.outputs, trainable = d.trainable, init = Seq(d.initializer), reg = d.reg)
    case b: Bias  => Leaf(Kind.Bias, units = b.features, trainable = b.trainable, init = Seq(Zeros, Zeros, b.initializer), reg = b.reg)
    case Activate(a) =>
      val (code, thr) = act(a)
      Leaf(Kind.Activate, activation = code, reluThreshold = thr)
    case Flatten => Leaf(Kind.Flatten)
    case c: Conv2D =>
      require(c.format == NHWC, "NCHW is outside the accelerated path (TensorFlow's CPU kernels do not support it either)")
      val (pad, explicit) = c.padding match {
        case Padding.Valid          => (Pad.Valid, (0, 0, 0, 0))
        case Padding.Same           => (Pad.Same, (0, 0, 0, 0))
        case Padding.Explicit(h, w) => (Pad.Explicit, (h._1, h._2, w._1, w._2))
      }
      Leaf(Kind.Conv2D, units = c.filters, window = c.kernel, strides = c.strides, padding = pad, explicitPad = explicit,
        trainable = c.trainable, init = Seq(c.initializer))
    case p: Pool2D =>
      require(p.format == NHWC, "NCHW is outside the accelerated path")
      val pad = p.padding match {
        case Padding.Valid => Pad.Valid
        case Padding.Same  => Pad.Same
        case _             => throw new IllegalArgumentException("requirement failed: only [Same, Valid] paddings are supported")
      }
      // the reference layer never forwards `reduce` (layer/Pool2D.scala:36-42): pool_honor_reduce stays 0
      Leaf(Kind.Pool2D, window = p.window, strides = p.strides, padding = pad, poolReduce = if (p.reduce == Reduce.Avg) 1 else 0)
    case bn: BatchNorm =>
      require(bn.axes == Seq(-1), "BatchNorm over axes other than the last one is outside the accelerated path")
      Leaf(Kind.BatchNorm, trainable = bn.trainable, bnMomentum = bn.momentum, bnEpsilon = bn.epsilon,
        bnFused = bn.fused.map(f => if (f) 1 else 0).getOrElse(-1),
        init = Seq(bn.gammaInitializer, bn.betaInitializer, bn.movingMeanInitializer, bn.movingVarianceInitializer))
    case RNN(cell: LSTMCell, returnSequence, stateful) =>
      val (a, _) = act(cell.activation)
      val (ra, _) = act(cell.recurrentActivation)
      Leaf(Kind.LSTM, units = cell.units, activation = a, recurrentActivation = ra, useBias = cell.bias, trainable = cell.trainable,
        init = Seq(cell.kernelInitializer, cell.recurrentInitializer, cell.biasInitializer, cell.biasForgetInitializer),
        returnSequence = returnSequence, stateful = stateful)
    case RNN(cell, returnSequence, stateful) =>
      // SimpleRNNCell(...) ?>> Bias ?>> activation (layer/RNN.scala:140-163): a composed cell of one to three leaves
      val parts = leaves(cell)
      val core = parts.collectFirst { case c: SimpleRNNCell => c }
        .getOrElse(throw new UnsupportedOperationException(s"RNN cell $cell is outside the accelerated path"))
      val bias = parts.collectFirst { case b: Bias => b }
      val (a, _) = parts.collectFirst { case Activate(x) => act(x) }.getOrElse((Act.Identity, 0f))
      Leaf(Kind.SimpleRNN, units = core.units, activation = a, useBias = bias.isDefined, trainable = core.trainable,
        init = Seq(core.kernelInitializer, core.recurrentInitializer, bias.map(_.initializer).getOrElse(Zeros)),
        returnSequence = returnSequence, stateful = stateful)
    case other => throw new UnsupportedOperationException(s"layer $other is outside the accelerated path")
  }

  private val random = new java.security.SecureRandom()

  /** sb_initializer {kind, has_seed, seed, a, b}.  `seed = None` draws one here: TF does the same per `initialize.eval`, and every
    * partition initialises independently on the first epoch (Optimizer.scala:123-131); the C ABI itself only takes explicit seeds. */
  private def writeInit(seg: MemorySegment, base: Long, in: Initializer): Unit = {
    def put(kind: Int, seed: Option[Long], a: Float = 0f, b: Float = 0f, random: Boolean = true): Unit = {
      seg.set(JAVA_INT, base, kind)
      seg.set(JAVA_INT, base + 4, if (random) 1 else 0)
      seg.set(JAVA_LONG, base + 8, if (random) seed.getOrElse(this.random.nextInt(Int.MaxValue).toLong) else 0L)
      seg.set(JAVA_FLOAT, base + 16, a)
      seg.set(JAVA_FLOAT, base + 20, b)
    }
    in match {
      case Zeros                               => put(Init.Zeros, None, random = false)
      case Ones                                => put(Init.Ones, None, random = false)
      case Const(v)                            => put(Init.Const, None, v, random = false)
      case RandomUniform(min, max, seed)       => put(Init.RandomUniform, seed, min, max)
      case RandomNormal(mean, std, seed)       => put(Init.RandomNormal, seed, mean, std)
      case RandomNormalTruncated(mean, std, s) => put(Init.TruncatedNormal, s, mean, std)
      case GlorotUniform(seed)                 => put(Init.GlorotUniform, seed)
      case GlorotNormal(seed)                  => put(Init.GlorotNormal, seed)
      case HeUniform(seed)                     => put(Init.HeUniform, seed)
      case HeNormal(seed)                      => put(Init.HeNormal, seed)
      case LecunUniform(seed)                  => put(Init.LecunUniform, seed)
      case LecunNormal(seed)                   => put(Init.LecunNormal, seed)
      case Orthogonal(gain, seed)              => put(Init.Orthogonal, seed, gain.getOrElse(0f))
      case other                               => throw new UnsupportedOperationException(s"initializer $other")
    }
  }

  private def offset(name: String): Long =
    Native.LayerDesc.byteOffset(MemoryLayout.PathElement.groupElement(name))

  private def writeLeaf(seg: MemorySegment, l: Leaf): Unit = {
    def int(name: String, v: Int): Unit = seg.set(JAVA_INT, offset(name), v)
    def flt(name: String, v: Float): Unit = seg.set(JAVA_FLOAT, offset(name), v)
    int("kind", l.kind); int("units", l.units); int("activation", l.activation); int("recurrent_activation", l.recurrentActivation)
    flt("relu_threshold", l.reluThreshold)
    int("window_h", l.window._1); int("window_w", l.window._2); int("stride_h", l.strides._1); int("stride_w", l.strides._2)
    int("padding", l.padding); int("pool_reduce", l.poolReduce); int("use_bias", if (l.useBias) 1 else 0)
    int("trainable", if (l.trainable) 1 else 0)
    flt("bn_momentum", l.bnMomentum); flt("bn_epsilon", l.bnEpsilon); int("bn_mode", 0 /* reference wiring */ ); int("bn_fused", l.bnFused)
    val initBase = offset("init")
    val initSize = Native.Initializer.byteSize()
    (0 until 4).foreach { slot =>
      writeInit(seg, initBase + slot * initSize, l.init.lift(slot).getOrElse(Zeros))
    }
    val (regKind, lambda) = l.reg match {
      case Zero       => (0, 0f)
      case L1(lambda) => (1, lambda)
      case L2(lambda) => (2, lambda)
      case other      => throw new UnsupportedOperationException(s"regularization $other")
    }
    int("reg_kind", regKind); flt("reg_lambda", lambda)
    int("pad_top", l.explicitPad._1); int("pad_bottom", l.explicitPad._2); int("pad_left", l.explicitPad._3); int("pad_right", l.explicitPad._4)
    int("return_sequence", if (l.returnSequence) 1 else 0); int("pool_honor_reduce", 0); int("stateful", if (l.stateful) 1 else 0)
  }

  private def lossCode(loss: Loss): Int = loss match {
    case MeanSquaredError        => 0
    case BinaryCrossentropy      => 1
    case CategoricalCrossentropy => 2
    case other                   => throw new UnsupportedOperationException(s"loss $other is outside the accelerated path")
  }

  /** (optimizer kind, rate, beta1, beta2, epsilon, init_acc, nesterov) -- the packing of scanet3_b200/dsl.py */
  private def algorithm(alg: Algorithm): (Int, Float, Float, Float, Float, Float, Boolean) = alg match {
    case SGD(rate, momentum, nesterov)             => (0, rate, momentum, 0f, 0f, 0f, nesterov)
    case Adam(rate, b1, b2, eps)                   => (1, rate, b1, b2, eps, 0f, false)
    case AdaGrad(rate, eps)                        => (2, rate, 0f, 0f, eps, 0f, false)
    case RMSProp(rate, rho, eps)                   => (3, rate, rho, 0f, eps, 0f, false)
    case AdaDelta(rate, rho, eps)                  => (4, rate, rho, 0f, eps, 0f, false)
    case Adamax(rate, b1, b2, eps, initAcc)        => (5, rate, b1, b2, eps, initAcc, false)
    case Nadam(b1, b2, eps, initAcc)               => (6, 0f, b1, b2, eps, initAcc, false)
    case AMSGrad(rate, b1, b2, eps)                => (7, rate, b1, b2, eps, 0f, false)
    case other                                     => throw new UnsupportedOperationException(s"algorithm $other")
  }

  /** Builds (sb_model_desc*, sb_train_desc*) in `arena` for `model` trained on batches shaped like `shard` (features, labels). */
  def of[A: Floating](
      model: Model,
      loss: Loss,
      alg: Algorithm,
      batchSize: Int,
      shard: (Tensor[A], Tensor[A]),
      precision: Int,
      arena: Arena,
      minimizing: Boolean = true): (MemorySegment, MemorySegment) = {
    val leafs = leaves(model).map(lower)
    val layerSize = Native.LayerDesc.byteSize()
    val layers = arena.allocate(layerSize * leafs.size, 8)
    leafs.zipWithIndex.foreach { case (leaf, i) => writeLeaf(layers.asSlice(i * layerSize, layerSize), leaf) }

    def dims(shape: Shape): Seq[Long] = shape.dims.tail.map(_.toLong) // per-sample shape: the row dimension is dropped
    val md = arena.allocate(Native.ModelDesc)
    def mdOff(name: String) = Native.ModelDesc.byteOffset(MemoryLayout.PathElement.groupElement(name))
    md.set(JAVA_INT, mdOff("n_layers"), leafs.size)
    md.set(ADDRESS, mdOff("layers"), layers)
    val in = dims(shard._1.shape)
    val lab = dims(shard._2.shape)
    md.set(JAVA_INT, mdOff("input_rank"), in.size)
    in.zipWithIndex.foreach { case (d, i) => md.set(JAVA_LONG, mdOff("input_dims") + 8L * i, d) }
    md.set(JAVA_INT, mdOff("label_rank"), lab.size)
    lab.zipWithIndex.foreach { case (d, i) => md.set(JAVA_LONG, mdOff("label_dims") + 8L * i, d) }

    val (kind, rate, b1, b2, eps, initAcc, nesterov) = algorithm(alg)
    val td = arena.allocate(Native.TrainDesc)
    def tdOff(name: String) = Native.TrainDesc.byteOffset(MemoryLayout.PathElement.groupElement(name))
    td.set(JAVA_INT, tdOff("batch"), batchSize)
    td.set(JAVA_INT, tdOff("loss"), lossCode(loss))
    td.set(JAVA_INT, tdOff("optimizer"), kind)
    td.set(JAVA_FLOAT, tdOff("rate"), rate)
    td.set(JAVA_FLOAT, tdOff("beta1"), b1)
    td.set(JAVA_FLOAT, tdOff("beta2"), b2)
    td.set(JAVA_FLOAT, tdOff("epsilon"), eps)
    td.set(JAVA_FLOAT, tdOff("init_acc"), initAcc)
    td.set(JAVA_INT, tdOff("nesterov"), if (nesterov) 1 else 0)
    td.set(JAVA_INT, tdOff("minimizing"), if (minimizing) 1 else 0)
    td.set(JAVA_INT, tdOff("precision"), precision)
    td.set(JAVA_INT, tdOff("use_graph"), 1)
    (md, td)
  }
}
