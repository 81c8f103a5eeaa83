

# Modules of this package that the drop-in does not provide (dataset_manilupation, dnn_models, ... / arg_extractor,
# data_provider, ...) resolve to the reference's own copies when its checkout follows the drop-in directory on
# sys.path: the package path is extended with every later directory of the same name.  The drop-in's modules come
# first, so `spotlight.interactions`, `spotlight.layers`, ... imported from those reference modules are the CUDA-backed
# ones.
from pkgutil import extend_path
__path__ = extend_path(__path__, __name__)
