"""Drop-in for the reference's module/weight_methods.py facade (`WeightMethods`, :727-761) restricted to the
method MTD-GAN trains with (`pcgrad`, :409-468); `from module.weight_methods import WeightMethods`
(train.py:18) resolves here.  No cvxpy/scipy import."""
from mtdgan_b200.weight_methods import WeightMethods, PCGrad, WeightMethod, METHODS  # noqa: F401

__all__ = ["WeightMethods", "PCGrad", "WeightMethod", "METHODS"]
