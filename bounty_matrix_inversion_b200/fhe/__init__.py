"""Drop-in for the part of `concrete.fhe` the reference uses."""
from . import tracing
from .compiler import Circuit, Compiler, Configuration, EncryptedData, PublicArguments, PublicResult
from .tracing import Tracer, one, ones, univariate, zero, zeros

# reference: `Tracer = fhe.tracing.tracer.Tracer` (qfloat.py:11, qfloat_matrix_inversion.py:17)
tracing.tracer = tracing


def compiler(parameter_encryption_statuses):
    """@fhe.compiler({...}) decorator form"""
    def wrap(function):
        return Compiler(function, parameter_encryption_statuses)
    return wrap


__all__ = ["Circuit", "Compiler", "Configuration", "EncryptedData", "PublicArguments", "PublicResult", "Tracer",
           "compiler", "one", "ones", "tracing", "univariate", "zero", "zeros"]
