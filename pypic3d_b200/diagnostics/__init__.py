"""I/O-boundary helpers of PyPIC3D/diagnostics used as parity observables: tile assembly, particle
flattening, and the `data/*.txt` row format.  Not on the step: plain tensor slicing, no kernels."""
from .output_adapters import (ParticleOutputRecord, assemble_tiled_scalar_field, assemble_tiled_vector_field,
                              scalar_field_for_output, vector_field_for_output, fields_for_output, particles_for_output)
from .plotting import write_data
