// Drop-in twin of scanet.optimizers.Optimizer.run (reference: src/main/scala/scanet/optimizers/Optimizer.scala:59-104) that
// keeps the reference's driver loop -- epochs, effects, stop condition, Step bookkeeping -- and replaces the three closures
// it calls (`backprop`, `loss`, `agg`, :132-134,213-227) with the C ABI of include/scanet_b200.h:
//   sc.broadcast(params)                  -> params stay resident on the k GPUs (set once at epoch 0)
//   mapPartitions(optimizeOnPartition)    -> k threads, one sb_engine per GPU, sb_train_epoch on the resident shard
//   treeReduce(aggregateParams)           -> sb_group_average (one fused NVLink pass; SPARK_CHAIN keeps the pairwise order)
// NOT COMPILED in this repository's CI (no JVM in the build image); the same loop is exercised through
// scanet3_b200/optimizer.py, which is a line-for-line Python mirror of this file.
package scanet.b200

import java.lang.foreign._
import java.lang.foreign.ValueLayout._
import java.util.concurrent.Executors

import scala.annotation.tailrec
import scala.concurrent.duration.Duration
import scala.concurrent.{Await, ExecutionContext, Future}

import scanet.core.{Floating, Params, Path, Shape, Tensor}
import scanet.models.{Loss, Model, TrainedModel}
import scanet.optimizers.Effect.StepContext // same effects / conditions as the reference
import scanet.optimizers.{Algorithm, Condition, Effect, Step, StepResult}

final case class B200Optimizer[A: Floating](
    alg: Algorithm,
    model: Model,
    loss: Loss,
    initParams: () => Option[Params[Tensor[A]]],
    shards: Seq[(Tensor[A], Tensor[A])], // partition i = contiguous rows, already (features, labels) tensors
    batchSize: Int,
    stop: Condition[StepContext[A]],
    doOnEach: Seq[Effect[A]],
    avgMode: Int = 1 /* SB_AVG_SPARK_CHAIN */,
    precision: Int = 1 /* SB_PREC_TF32 */ ) {

  private val lossModel = model.withLoss(loss)
  private val k = shards.size

  def run(): TrainedModel[A] = {
    // Optimizer.scala:65 — the reference's own `Model.describe` still runs on the JVM side
    println(s"Training model:\n${model.describe[A](batchSize +: shards.head._1.shape.tail)}")
    val pool = Executors.newFixedThreadPool(k) // one JVM thread per GPU, like one Spark task per partition
    implicit val ec: ExecutionContext = ExecutionContext.fromExecutor(pool)
    val session = Arena.ofShared()
    try {
      // ---- compileBackprop / compileLoss (:163-194) -> one engine per GPU ----
      val (modelDesc, trainDesc) = Descriptors.of(model, loss, alg, batchSize, shards.head, precision, session)
      val engines = (0 until k).map { dev =>
        val out = session.allocate(ADDRESS)
        Native.check(Native.engineCreate.invoke(modelDesc, trainDesc, dev, out).asInstanceOf[Int])
        out.get(ADDRESS, 0)
      }
      // ---- epoch 0 initialisation (:123-131) ----
      engines.foreach { e =>
        initParams() match {
          case Some(ps) =>
            ps.params.foreach { case (path, t) => setParam(e, path, t, session) }
            Native.check(Native.initOptParams.invoke(e).asInstanceOf[Int])
          case None => Native.check(Native.initParams.invoke(e).asInstanceOf[Int])
        }
      }
      // ---- partition iterator + TensorIterator (Iterators.scala:39-83) -> shard resident in HBM ----
      engines.zip(shards).foreach { case (e, (x, y)) =>
        Native.check(Native.datasetUpload
          .invoke(e, TensorBridge.segment(x), TensorBridge.segment(y), x.shape.head.toLong).asInstanceOf[Int])
      }
      val group = {
        val arr = session.allocate(ADDRESS, k.toLong)
        engines.zipWithIndex.foreach { case (e, i) => arr.setAtIndex(ADDRESS, i.toLong, e) }
        val out = session.allocate(ADDRESS)
        Native.check(Native.groupCreate.invoke(arr, k, avgMode, MemorySegment.NULL, 0 /* P2P */, out).asInstanceOf[Int])
        out.get(ADDRESS, 0)
      }

      @tailrec
      def optimize(prevStep: Step[A]): Unit = {
        val start = System.currentTimeMillis()
        // mapPartitions(optimizeOnPartition): all full batches of every shard, iter = globalIter + j + 1 (:135-157)
        val perRank = Await.result(Future.sequence(engines.map { e =>
          Future {
            val iters = session.allocate(JAVA_LONG)
            val lossOut = session.allocate(JAVA_FLOAT)
            Native.check(Native.trainEpoch.invoke(e, prevStep.iter.toLong, iters, lossOut).asInstanceOf[Int])
            (iters.get(JAVA_LONG, 0), lossOut.get(JAVA_FLOAT, 0))
          }
        }), Duration.Inf)
        // treeReduce(aggregateParams) (:86,196-230): params + optimizer state averaged in place, loss combined with the
        // same weights, iterations summed
        val losses = session.allocate(JAVA_FLOAT, k.toLong)
        val iters = session.allocate(JAVA_LONG, k.toLong)
        perRank.zipWithIndex.foreach { case ((n, l), i) =>
          iters.setAtIndex(JAVA_LONG, i.toLong, n); losses.setAtIndex(JAVA_FLOAT, i.toLong, l)
        }
        Native.check(Native.groupAverage.invoke(group, losses, iters).asInstanceOf[Int])
        val finish = System.currentTimeMillis()
        val iterations = iters.getAtIndex(JAVA_LONG, 0).toInt
        val step: Step[A] = prevStep.nextEpoch.incIter(iterations)
        // effects read params lazily: nothing leaves the device unless an effect asks (RecordAccuracy, :52-64)
        val result = TensorBridge.lazyStepResult[A](engines.head, iterations, losses.getAtIndex(JAVA_FLOAT, 0), model, alg, batchSize)
        val stepCtx = StepContext(step, result, lossModel, finish - start)
        doOnEach.foldLeft(Effect.State(Effect.Console(), null))((ctx, effect) => effect(ctx, stepCtx)).run()
        if (!stop(stepCtx)) optimize(step)
      }
      optimize(Step(batchSize))
      val params = TensorBridge.readModelParams[A](engines.head, model, batchSize) // sb_params_get -> Tensor.fromBytes
      Native.groupDestroy.invoke(group)
      engines.foreach(e => Native.engineDestroy.invoke(e))
      lossModel.trained(params)
    } finally {
      pool.shutdown()
      session.close()
    }
  }

  private def setParam(e: MemorySegment, path: Path, t: Tensor[A], session: Arena): Unit =
    Native.check(Native.paramsSet
      .invoke(e, session.allocateFrom(path.toString), TensorBridge.segment(t), t.shape.power.toLong).asInstanceOf[Int])
}
