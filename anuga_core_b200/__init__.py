"""anuga_core_b200 - B200-native backend for ANUGA's discontinuous-elevation
shallow-water timestep (DE0/DE1/DE2): hand-written sm_100a CUDA behind the
reference's shallow_water.Domain API.  See DESIGN.md and INTEGRATION.md."""
from .mesh import Mesh, rectangular_cross, rectangular, morton_order
from .quantity import Quantity
from .boundaries import (Characteristic_stage_boundary, Dirichlet_discharge_boundary, Flather_external_stage_zero_velocity_boundary, Reflective_boundary, Dirichlet_boundary, Transmissive_boundary, Time_boundary,
                         Transmissive_n_momentum_zero_t_momentum_set_stage_boundary,
                         Transmissive_momentum_set_stage_boundary,
                         Transmissive_stage_zero_momentum_boundary, Time_stage_zero_momentum_boundary)
from .operators import (Rate_operator, Set_quantity, Set_stage, Set_quantity_operator,
                        Set_stage_operator, Set_elevation, Set_elevation_operator)
from .structures import (Region, Inlet, Inlet_operator, Inlet_enquiry, Structure_operator,
                         Boyd_box_operator, Boyd_pipe_operator, Weir_orifice_trapezoid_operator,
                         Internal_boundary_operator, pumping_station_function)
from .forcing import Wind_stress, General_forcing, Rainfall, Inflow
from .file_boundary import File_boundary, Field_boundary, Time_space_boundary, file_function, timefile2netcdf
from .domain import Domain, rectangular_cross_domain, load_checkpoint_file, MODE_B200
from .backend import SwkError, device_count
from .attach import set_multiprocessor_mode_b200, B200_interface
# the reference's script-level parallel API (anuga.distribute, myid, numprocs, barrier, finalize)
from .parallel import distribute_collective as distribute, myid, numprocs, barrier, finalize

from .mesh_io import create_domain_from_file, read_tsh
from .compat import install_as_anuga, Polygon_function, read_polygon, inside_polygon, g, indent

__version__ = "0.1.0"
