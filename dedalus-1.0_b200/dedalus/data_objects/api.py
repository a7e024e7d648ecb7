from .fields import create_field_classes, BaseField, ScalarFieldBase, VectorFieldBase, TensorFieldBase
from .state_data import StateData
from .representations import Representation, FourierRepresentation, FourierShearRepresentation, ChebyshevRepresentation
from .aux_equation import AuxEquation
