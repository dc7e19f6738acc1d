"""Optimizer handle mirroring ``tf.keras.optimizers.Adam`` as the reference uses it
(train.py:99-101; camera_streamer_qt.py:586-588,1329).  The update itself is the fused
CUDA Adam kernel inside libkcvae.so (Keras optimizer_v2 formula, epsilon 1e-7)."""


class Adam:
    def __init__(self, learning_rate=0.001, beta_1=0.9, beta_2=0.999, epsilon=1e-7, amsgrad=False, name="Adam"):
        if amsgrad:
            raise NotImplementedError("amsgrad is not used by the reference")
        if (beta_1, beta_2, epsilon) != (0.9, 0.999, 1e-7):
            raise NotImplementedError("only the Keras default beta_1/beta_2/epsilon are implemented")
        self.learning_rate = float(learning_rate)   # mutable at run time, read before every step
        self.beta_1, self.beta_2, self.epsilon = beta_1, beta_2, epsilon
        self.iterations = 0
        self.name = name

    @property
    def lr(self):
        return self.learning_rate

    @lr.setter
    def lr(self, v):
        self.learning_rate = float(v)

    def get_config(self):
        return {"name": self.name, "learning_rate": self.learning_rate, "beta_1": self.beta_1,
                "beta_2": self.beta_2, "epsilon": self.epsilon, "amsgrad": False}
