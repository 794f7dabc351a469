"""scanet3_b200: B200-native engine for scanet3's data-parallel train path (per-worker minibatch step +
epoch-end model averaging).  Host-side mirror of the reference's DSL over `libscanet_b200.so`."""
from . import _native as native
from ._native import (AVG_MEAN, AVG_SPARK_CHAIN, AVG_WEIGHTS, COLL_NCCL, COLL_P2P, PREC_3XTF32, PREC_FP32, PREC_TF32,
                      IllegalArgument, NativeError, Unsupported)
from .dsl import *  # noqa: F401,F403
from .dsl import (Explicit, Activate, Activation, AdaDelta, AdaGrad, Adam, Adamax, Algorithm, AMSGrad, BatchNorm, Bias,
                  BinaryCrossentropy, CategoricalCrossentropy, Composed, Const, Conv2D, Dense, Flatten, GlorotNormal, GlorotUniform, HeNormal,
                  HeUniform, Identity, Initializer, L1, L2, LecunNormal, LecunUniform, LinearRegression, LogisticRegression, Loss, LSTM,
                  LSTMCell, MeanSquaredError, Nadam, Ones, Orthogonal, Pool2D, RandomNormal, RandomNormalTruncated, RandomUniform, ReLU, RMSProp,
                  RNN, Same, SGD, Sigmoid, SimpleRNN, SimpleRNNCell, Softmax, Tanh, Valid, Zero, Zeros)
from .engine import Engine, Group, spark_chain_weights, wire_decode, wire_encode
from .optimizer import (Builder, Dataset, LossModel, Optimizer, RecordAccuracy, RecordIteration, RecordLoss, SaveCheckpoint, Step,
                        TrainedModel, MAE, MSE, R2Score, RMSE, accuracy, always, epochs, iterations, maximize, minimize, never)
