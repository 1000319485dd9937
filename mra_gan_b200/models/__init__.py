"""Model registry with the reference's plugin convention (models/__init__.py:4-43 of the
reference): ``create_model(opt)`` imports ``<opt.model>_model`` from this package and instantiates
the BaseModel subclass whose lower-cased name is ``<model without underscores>model``."""
import importlib

from .base_model import BaseModel


def find_model_using_name(model_name):
    module = importlib.import_module("%s.%s_model" % (__name__, model_name))
    wanted = model_name.replace("_", "").lower() + "model"
    found = [cls for name, cls in vars(module).items()
             if isinstance(cls, type) and issubclass(cls, BaseModel) and name.lower() == wanted]
    if not found:
        raise ImportError("%s_model.py defines no BaseModel subclass named like %r" % (model_name, wanted))
    return found[-1]


def get_option_setter(model_name):
    return find_model_using_name(model_name).modify_commandline_options


def create_model(opt):
    instance = find_model_using_name(opt.model)()
    instance.initialize(opt)
    print("model [%s] was created" % instance.name())
    return instance
