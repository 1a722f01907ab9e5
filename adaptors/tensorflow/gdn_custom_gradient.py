"""What a reference maintainer would put in GDN.call (tensorflow_compression/python/layers/gdn.py:371-421) to run
the layer on libtfcb200.so where the exponents are fixed (the default alpha = epsilon = 1 of bls2017 / bmshj2018;
also alpha = 2, epsilon = 1/2).  Not imported by this repository (no TensorFlow in its image)."""
import tensorflow as tf

_ops = tf.load_op_library("libtfcb200_tf.so")  # built from tfcb200_tf_gpu_kernels.cc, see README.md


def gdn_b200(inputs, gamma, beta, inverse=False, rectify=False, alpha=1.0, epsilon=1.0):
  """Channels-last GDN / IGDN with the hand-derived backward pass of the CUDA library."""
  kw = dict(inverse=inverse, rectify=rectify, alpha=float(alpha), epsilon=float(epsilon))

  @tf.custom_gradient
  def fn(x, g, b):
    y = _ops.gdn_forward(x, g, b, **kw)

    def grad(dy):
      return _ops.gdn_backward(x, g, b, dy, **kw)  # (dx, dgamma, dbeta)

    return y, grad

  return fn(tf.cast(inputs, tf.float32), gamma, beta)
