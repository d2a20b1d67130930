from .spec import ConvNetSpec, MLPSpec
from .mnist_conv_nn import MNISTConvNet
from .fourier_nn import FourierNet, SIRENLayer
from .relu_nn import FFReLUNet, FFTanhNet, FFSigmoidNet

__all__ = ["ConvNetSpec", "MLPSpec", "MNISTConvNet", "FourierNet", "SIRENLayer",
           "FFReLUNet", "FFTanhNet", "FFSigmoidNet"]
