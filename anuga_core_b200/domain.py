"""shallow_water.Domain API on top of the device backend.

Host-side mirror of the reference interface for the DE hot path: the same
method names, argument meaning and error behaviour as

  anuga/shallow_water/shallow_water_domain.py   class Domain :144
      set_flow_algorithm :1311-1400 (presets DE0 :583, DE1 :646, DE2 :709, DE0_7 :835, DE1_7 :770)
      set_quantity :907, evolve :2300-2407, set_multiprocessor_mode :2859-2899
  anuga/abstract_2d_finite_volumes/generic_domain.py
      set_boundary :937-1033, _evolve_base :1715-1912, evolve_one_*_step :1914-2179,
      update_timestep :2349-2415, set_fixed_flux_timestep :2332

The arrays of the conserved quantities live in HBM between yields; the numpy
arrays on ``domain.quantities[...]`` are refreshed at every yield (centroid
values eagerly, edge/vertex values on first access) and re-uploaded when the
user changes them through ``set_quantity``/``set_values`` (or calls
``sync_from_host``).  There is no CPU execution path: without a usable sm_100
device ``evolve`` raises.
"""
import numpy as np

from . import backend as _b
from .mesh import Mesh, rectangular_cross, rectangular_cross_neighbours, morton_order
from .quantity import Quantity

# multiprocessor_mode of the B200 backend.  The reference uses 0 (orig C), 1 (simd),
# 2 (openmp), 3 (openacc), 4 (cupy experiment) - shallow_water_domain.py:2859-2876.
MODE_B200 = 5

_CONFIG = dict(
    epsilon=1.0e-12, g=9.8, minimum_allowed_height=1.0e-05, maximum_allowed_speed=0.0,
    max_timestep=1000.0, min_timestep=1.0e-6, max_smallsteps=50, low_froude=0,
    extrapolate_velocity_second_order=True, sloped_mannings_function=False,
    optimise_dry_cells=False, default_order=2,
)

_ALGORITHMS = {
    #          CFL  method   min_allowed_height  beta  beta_dry
    "DE0":   (0.9, "euler", 1.0e-12, 0.5, 0.0),
    "DE1":   (1.0, "rk2",   1.0e-5,  1.0, 0.0),
    "DE2":   (1.0, "rk3",   1.0e-5,  1.0, 0.0),
    "DE0_7": (0.9, "euler", 1.0e-12, 0.7, 0.1),
    "DE1_7": (1.0, "rk2",   1.0e-12, 0.75, 0.1),
}


class Domain:
    conserved_quantities = ["stage", "xmomentum", "ymomentum"]
    evolved_quantities = ["stage", "xmomentum", "ymomentum", "elevation", "height", "xvelocity", "yvelocity"]
    other_quantities = ["elevation", "friction", "height", "xvelocity", "yvelocity", "x", "y"]

    def __init__(self, coordinates=None, vertices=None, boundary=None, mesh=None,
                 use_inscribed_circle=False, full_send_dict=None, ghost_recv_dict=None,
                 processor=0, numproc=1, number_of_full_nodes=None, number_of_full_triangles=None,
                 ghost_layer_width=2, verbose=False, device=0, reorder=True, **ignored):
        if mesh is None:
            mesh = Mesh(coordinates, vertices, boundary, use_inscribed_circle=use_inscribed_circle)
        self.mesh = mesh
        for name in ("nodes", "triangles", "neighbours", "neighbour_edges", "surrogate_neighbours",
                     "number_of_boundaries", "normals", "edgelengths", "radii", "areas",
                     "centroid_coordinates", "vertex_coordinates", "edge_midpoint_coordinates",
                     "boundary", "boundary_cells", "boundary_edges", "tag_boundary_cells",
                     "number_of_triangles", "number_of_nodes", "boundary_length"):
            setattr(self, name, getattr(mesh, name))
        self.edge_coordinates = self.edge_midpoint_coordinates
        self.number_of_elements = self.number_of_triangles
        N = self.number_of_triangles
        self.verbose = verbose
        self.device = device
        self.reorder = reorder
        self.processor = processor
        self.numproc = numproc
        self.ghost_layer_width = ghost_layer_width
        self.full_send_dict = {} if full_send_dict is None else full_send_dict
        self.ghost_recv_dict = {} if ghost_recv_dict is None else ghost_recv_dict
        # generic_domain.py:285-291
        self.tri_full_flag = np.ones(N, dtype=np.int64)
        for key in self.ghost_recv_dict:
            self.tri_full_flag[np.asarray(self.ghost_recv_dict[key][0], dtype=np.int64)] = 0
        self.number_of_full_triangles = int(self.tri_full_flag.sum()) if number_of_full_triangles is None \
            else number_of_full_triangles
        self.number_of_full_nodes = self.number_of_nodes if number_of_full_nodes is None else number_of_full_nodes

        self.quantities = {}
        for name in ["stage", "xmomentum", "ymomentum", "elevation", "friction", "height", "xvelocity", "yvelocity"]:
            self.quantities[name] = Quantity(self, name)
            self.quantities[name]._fetch = self._fetch_lazy

        # scalars (config.py) and the DE0 default preset (shallow_water_domain.py:319-334)
        self.epsilon = _CONFIG["epsilon"]
        self.g = _CONFIG["g"]
        self.H0 = _CONFIG["minimum_allowed_height"]
        self.maximum_allowed_speed = _CONFIG["maximum_allowed_speed"]
        self.evolve_max_timestep = _CONFIG["max_timestep"]
        self.evolve_min_timestep = _CONFIG["min_timestep"]
        self.max_smallsteps = _CONFIG["max_smallsteps"]
        self.low_froude = _CONFIG["low_froude"]
        self.extrapolate_velocity_second_order = _CONFIG["extrapolate_velocity_second_order"]
        self.use_sloped_mannings = _CONFIG["sloped_mannings_function"]
        self.optimise_dry_cells = _CONFIG["optimise_dry_cells"]
        self.default_order = _CONFIG["default_order"]
        self._order_ = self.default_order
        self.centroid_transmissive_bc = False
        self.fixed_flux_timestep = None
        self.set_flow_algorithm("DE0")

        self.multiprocessor_mode = MODE_B200
        self.boundary_map = None
        self.fractional_step_operators = []
        from .riverwall import RiverWall
        self.riverwallData = RiverWall(self)                # shallow_water_domain.py:402
        from .forcing import manning_friction_implicit
        self.forcing_terms = [manning_friction_implicit]     # shallow_water_domain.py:300
        self._forcing_on_device = False
        self.starttime = 0.0
        self.relative_time = 0.0
        self.evolve_starttime = 0.0
        self.timestep = 0.0
        self.flux_timestep = 0.0
        self.yieldstep = None
        self.finaltime = None
        self.relative_finaltime = None
        self.evolved_called = False
        self.number_of_steps = 0
        self.number_of_first_order_steps = 0
        self.recorded_min_timestep = self.evolve_max_timestep
        self.recorded_max_timestep = self.evolve_min_timestep
        self.smallsteps = 0
        self.boundary_flux_integral = 0.0
        self.fractional_step_volume_integral = 0.0
        self.total_steps = 0
        self.kernel_launches = 0
        self.store = False                        # the reference defaults to True; SWW output is opt-in here
        self.name = "domain"
        self.datadir = "."
        self.smooth = True                        # set_store_vertices_uniquely(False), :341
        self.store_centroids = True               # _set_DE*_defaults
        self.using_centroid_averaging = True      # _set_DE*_defaults (shallow_water_domain.py:596)
        self.minimum_storable_height = 1.0e-3     # anuga/config.py:189
        self.quantities_to_be_stored = {"elevation": 1, "friction": 1, "stage": 2, "xmomentum": 2, "ymomentum": 2}
        self.yieldstep_counter = 0
        self.output_frequency = 1
        self.writer = None
        self.checkpoint = False
        self._dev = None
        self.pin_host_arrays = True
        self._stale = set()
        self.timestep_history = []
        self.record_timestep_history = False

    def __len__(self):
        return self.number_of_triangles

    # ------------------------------------------------------------------
    # configuration (same names as the reference)
    # ------------------------------------------------------------------
    def set_flow_algorithm(self, flag="DE0"):
        flag = str(flag).replace(".", "_")
        if flag not in _ALGORITHMS:
            raise Exception("Flow algorithm %r is not part of the B200 hot path; supported: %s"
                            % (flag, sorted(_ALGORITHMS)))
        cfl, method, mah, beta, beta_dry = _ALGORITHMS[flag]
        self.flow_algorithm = flag
        # _set_config_defaults re-reads config.py (shallow_water_domain.py:486-533)
        self.H0 = _CONFIG["minimum_allowed_height"]
        self.g = _CONFIG["g"]
        self.low_froude = _CONFIG["low_froude"]
        self.use_sloped_mannings = _CONFIG["sloped_mannings_function"]
        self.CFL = cfl
        self.timestepping_method = method
        self.minimum_allowed_height = mah
        self.default_order = 2
        self.extrapolate_velocity_second_order = True
        self.beta_w = self.beta_uh = self.beta_vh = beta
        self.beta_w_dry = self.beta_uh_dry = self.beta_vh_dry = beta_dry
        self.optimise_dry_cells = False
        self.maximum_allowed_speed = 0.0
        self._params_dirty = True

    def get_flow_algorithm(self):
        return self.flow_algorithm

    def get_using_discontinuous_elevation(self):
        return True

    def set_timestepping_method(self, flag):
        m = {1: "euler", 2: "rk2", 3: "rk3", "euler": "euler", "rk2": "rk2", "rk3": "rk3"}
        if flag not in m:
            raise Exception("Incorrect option for set_timestepping_method")
        self.timestepping_method = m[flag]
        self._params_dirty = True

    def get_timestepping_method(self):
        return self.timestepping_method

    def set_CFL(self, cfl=1.0):
        """generic_domain.py:1069-1081: warns above 2, must be positive"""
        if cfl > 2.0:
            import warnings
            warnings.warn("Setting CFL > 2.0")
        assert cfl > 0.0
        self.CFL = cfl
        self._params_dirty = True

    set_cfl = set_CFL

    def get_CFL(self):
        return self.CFL

    get_cfl = get_CFL

    def set_default_order(self, n):
        """spatial order 1 or 2 (generic_domain.py:956-962)"""
        assert n in [1, 2], "Default order must be either 1 or 2. I got %s" % n
        self.default_order = n
        self._order_ = self.default_order
        self._params_dirty = True

    def set_maximum_allowed_speed(self, maximum_allowed_speed):
        # (any value but 0 is refused when the device handle is created: not part of the DE path)
        self.maximum_allowed_speed = maximum_allowed_speed
        self._params_dirty = True

    def get_beta(self):
        return self.beta

    def get_minimum_allowed_height(self):
        return self.minimum_allowed_height

    def get_minimum_storable_height(self):
        return self.minimum_storable_height

    def get_evolve_max_timestep(self):
        return self.evolve_max_timestep

    def get_evolve_min_timestep(self):
        return self.evolve_min_timestep

    def get_centroid_transmissive_bc(self):
        return self.centroid_transmissive_bc

    def get_store_centroids(self):
        return self.store_centroids

    def get_algorithm_parameters(self):
        """the parameters of the DE path that are currently set (shallow_water_domain.py:1031-1063)"""
        return dict(minimum_allowed_height=self.minimum_allowed_height,
                    maximum_allowed_speed=self.maximum_allowed_speed,
                    minimum_storable_height=self.minimum_storable_height, g=self.g,
                    optimise_dry_cells=self.optimise_dry_cells, low_froude=self.low_froude,
                    use_sloped_mannings=self.use_sloped_mannings,
                    compute_fluxes_method="DE", distribute_to_vertices_and_edges_method="DE",
                    flow_algorithm=self.get_flow_algorithm(), CFL=self.get_CFL(),
                    timestepping_method=self.get_timestepping_method(),
                    extrapolate_velocity_second_order=self.extrapolate_velocity_second_order)

    def print_algorithm_parameters(self):
        print("#============================")
        print("# Domain Algorithm Parameters ")
        print("#============================")
        parameters = self.get_algorithm_parameters()
        for key in sorted(parameters.keys()):
            print("# %-41s:  %s" % (key, parameters[key]))
        print("#----------------------------")

    def set_beta(self, beta):
        self.beta = beta
        self.beta_w = self.beta_uh = self.beta_vh = beta
        self.beta_w_dry = self.beta_uh_dry = self.beta_vh_dry = beta
        self._params_dirty = True

    def set_betas(self, beta_w, beta_w_dry, beta_uh, beta_uh_dry, beta_vh, beta_vh_dry):
        self.beta_w, self.beta_w_dry = beta_w, beta_w_dry
        self.beta_uh, self.beta_uh_dry = beta_uh, beta_uh_dry
        self.beta_vh, self.beta_vh_dry = beta_vh, beta_vh_dry
        self._params_dirty = True

    def set_minimum_allowed_height(self, minimum_allowed_height):
        # shallow_water_domain.py:1476-1492: the only place H0 is set
        self.minimum_allowed_height = minimum_allowed_height
        self.H0 = minimum_allowed_height
        self._params_dirty = True

    def set_low_froude(self, low_froude=0):
        assert low_froude in (0, 1, 2)
        self.low_froude = low_froude
        self._params_dirty = True

    def set_extrapolate_velocity(self, flag=True):
        self.extrapolate_velocity_second_order = bool(flag)
        self._params_dirty = True

    def set_sloped_mannings_function(self, flag=True):
        self.use_sloped_mannings = bool(flag)
        self._params_dirty = True
        if self._dev is not None:
            self._release_device()

    def set_centroid_transmissive_bc(self, flag):
        self.centroid_transmissive_bc = bool(flag)
        self._params_dirty = True

    def set_fixed_flux_timestep(self, flux_timestep=None):
        if flux_timestep is not None and not flux_timestep > 0.0:
            raise Exception("flux_timestep needs to be greater than 0.0")
        self.fixed_flux_timestep = flux_timestep
        self._params_dirty = True

    def set_evolve_max_timestep(self, t):
        self.evolve_max_timestep = t
        self._params_dirty = True

    def set_evolve_min_timestep(self, t):
        self.evolve_min_timestep = t
        self._params_dirty = True

    def set_multiprocessor_mode(self, multiprocessor_mode=MODE_B200):
        """Only the device mode exists here and it never falls back to a CPU mode
        (the reference's mode 4 falls back to mode 0, shallow_water_domain.py:2887-2894)."""
        if multiprocessor_mode != MODE_B200:
            raise Exception("anuga_core_b200 implements multiprocessor_mode %d (B200 device) only; "
                            "modes 0-4 are the reference's CPU/CuPy backends" % MODE_B200)
        self.multiprocessor_mode = multiprocessor_mode
        self._ensure_device()

    def get_multiprocessor_mode(self):
        return self.multiprocessor_mode

    def compute_flux_update_frequency(self, *a, **k):
        """local time-stepping is a no-op in the reference's modes 2-4 as well
        (sw_domain_openmp_ext.pyx:413-415)"""
        return None

    # checkpointing (shallow_water_domain.py:2376-2397 pickles the whole Domain): bring the state to
    # the host and drop the device handle; it is rebuilt on the next evolve()
    def __getstate__(self):
        self.sync_to_host()
        state = dict(self.__dict__)
        state["_dev"] = None
        state["_comm"] = None
        state["_segments"] = {}
        state.pop("_pushed_once", None)
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)
        for q in self.quantities.values():
            q.host_dirty = True
            q._fetch = self._fetch_lazy
        self._stale = set()
        self._lazy_stale = set()
        for op in self.fractional_step_operators:
            if hasattr(op, "op_id"):
                op.op_id = None

    # -- yield-time SWW output (anuga_core_b200/sww.py) ------------------------------------
    def set_store(self, flag=True):
        self.store = bool(flag)

    def get_store(self):
        return self.store

    def set_name(self, name=None, timestamp=False):
        if name is None:
            name = "domain"
        if name.endswith(".sww"):
            name = name[:-4]
        if timestamp:
            from time import localtime, strftime
            name = name + "_" + strftime("%Y%m%d_%H%M%S", localtime())
        self.global_name = name
        if self.numproc > 1:          # Parallel_domain.set_name (parallel_shallow_water.py:110-120)
            name = name + "_P%d_%d" % (self.numproc, self.processor)
        self.name = name

    def get_name(self):
        return self.name

    def get_global_name(self):
        return getattr(self, "global_name", self.name)

    def sww_merge(self, verbose=False, delete_old=False):
        """one global SWW file from the per-rank files of a distributed run
        (parallel_shallow_water.py:158-182); a no-op for a sequential domain"""
        comm = getattr(self, "_comm", None)
        if comm is not None:
            comm.barrier()
        if self.processor == 0 and self.numproc > 1 and self.store:
            import os
            from .sww_merge import sww_merge_parallel
            sww_merge_parallel(os.path.join(self.get_datadir(), self.get_global_name()), self.numproc,
                               verbose, delete_old)
        if comm is not None:
            comm.barrier()

    def set_datadir(self, path):
        self.datadir = path

    def get_datadir(self):
        return self.datadir

    def set_store_centroids(self, flag=True):
        self.store_centroids = flag

    def set_store_vertices_uniquely(self, flag=True, reduction=None):
        self.smooth = not flag

    def set_store_vertices_smoothly(self, flag=True, reduction=None):
        self.smooth = flag

    def set_using_centroid_averaging(self, flag=True):
        self.using_centroid_averaging = bool(flag)

    def get_using_centroid_averaging(self):
        return self.using_centroid_averaging

    def set_minimum_storable_height(self, h):
        self.minimum_storable_height = h

    def set_checkpointing(self, checkpoint=True, checkpoint_dir="CHECKPOINTS", checkpoint_step=10,
                          checkpoint_time=None):
        """pickle the domain every `checkpoint_step` yields, or every `checkpoint_time` seconds of wall time
        (shallow_water_domain.py:1123-1157); load_checkpoint_file() picks the run up again"""
        if not checkpoint:
            self.checkpoint = False
            return
        import os
        import time
        os.makedirs(checkpoint_dir, exist_ok=True)
        self.checkpoint_dir = checkpoint_dir
        if checkpoint_time is not None:
            self.walltime_prev = time.time()
            self.checkpoint_time = checkpoint_time
            self.checkpoint_step = 0
        else:
            self.checkpoint_step = checkpoint_step
        self.checkpoint = True

    def _checkpoint_if_due(self):
        """the checkpoint block of Domain.evolve (shallow_water_domain.py:2376-2397)"""
        import os
        import time
        save = False
        if self.checkpoint_step == 0:
            save = time.time() - self.walltime_prev > self.checkpoint_time
            comm = getattr(self, "_comm", None)
            if comm is not None and comm.size > 1:      # rank 0 decides for everybody
                flag = comm.allreduce_max(float(save) if comm.rank == 0 else 0.0)
                save = flag > 0.0
        elif self.yieldstep_counter % self.checkpoint_step == 0:
            save = True
        if save:
            self.save_checkpoint()
            self.walltime_prev = time.time()

    def save_checkpoint(self):
        import os
        try:
            import dill as pickle
        except ImportError:
            import pickle
        name = os.path.join(self.checkpoint_dir, self.get_name()) + "_" + str(self.get_time()) + ".pickle"
        with open(name, "wb") as f:
            pickle.dump(self, f)
        comm = getattr(self, "_comm", None)
        if comm is not None:
            comm.barrier()
        return name

    def initialise_storage(self):
        """shallow_water_domain.py:2410-2423"""
        from .sww import SWW_file
        self.writer = SWW_file(self)
        self.writer.store_connectivity()

    def store_timestep(self):
        self.writer.store_timestep()

    def set_quantities_to_be_stored(self, q):
        self.quantities_to_be_stored = q

    def set_starttime(self, t):
        self.starttime = float(t)

    # ------------------------------------------------------------------
    # quantities and boundaries
    # ------------------------------------------------------------------
    def set_quantity(self, name, *args, **kwargs):
        if name not in self.quantities:
            raise Exception("Quantity %r is not part of this domain" % name)
        self.quantities[name].set_values(*args, **kwargs)
        self._stale.discard(name)

    def add_quantity(self, name, *args, **kwargs):
        """name += whatever set_quantity accepts (generic_domain.py:880-905)"""
        Q = Quantity(self)
        Q.set_values(*args, **kwargs)
        self.set_quantity(name, self.quantities[name] + Q)

    def create_quantity_from_expression(self, expression):
        """new Quantity from an arithmetic expression over the domain's quantities, e.g.
        'stage - elevation' (generic_domain.py:916-935; only names of quantities and numbers are
        visible to the expression)"""
        return eval(expression, {"__builtins__": {}}, dict(self.quantities))

    def get_quantity(self, name):
        return self.quantities[name]

    def get_quantity_names(self):
        return list(self.quantities.keys())

    def get_boundary_tags(self):
        return self.mesh.get_boundary_tags()

    def set_boundary(self, boundary_map):
        """generic_domain.py:937-1033: every tag of the mesh must be bound."""
        if self.boundary_map is None:
            self.boundary_map = dict(boundary_map)
        else:
            for key in boundary_map:
                self.boundary_map[key] = boundary_map[key]
        if self.numproc > 1 and "ghost" not in self.boundary_map:
            self.boundary_map["ghost"] = None       # outer edges of the ghost layer (parallel_api.py:129)
        for tag in self.get_boundary_tags():
            if tag not in self.boundary_map:
                raise Exception("Tag \"%s\" has not been bound to a boundary object.\n"
                                "All boundary tags defined in domain must appear in set_boundary.\n"
                                "The tags are: %s" % (tag, self.get_boundary_tags()))
        self.boundary_objects = [((int(v), int(e)), self.boundary_map[self.mesh.boundary[(v, e)]])
                                 for (v, e) in sorted(self.mesh.boundary.keys())]
        self._boundary_dirty = True

    def set_riverwall_tables(self, edge_flux_type, riverwall_elevation, hydraulic_properties_rowIndex,
                             hydraulic_properties):
        """Riverwall edge tables as structures/riverwall.py:406-415 leaves them on the domain:
        edge_flux_type (3N,) == 1 on wall edges, one elevation / row index per wall edge in (k, i)
        order, hydraulic_properties rows = [Qfactor, s1, s2, h1, h2].  (Building them from
        breaklines is mesh set-up, outside the hot path.)"""
        eft = np.ascontiguousarray(edge_flux_type, dtype=np.int64).reshape(-1)
        assert eft.size == 3 * self.number_of_triangles
        self.edge_flux_type = eft
        counter = np.zeros_like(eft)
        idx = np.flatnonzero(eft == 1)
        counter[idx] = np.arange(1, idx.size + 1)
        self.edge_river_wall_counter = counter
        self.number_of_riverwall_edges = int(idx.size)
        self.riverwall_elevation = np.ascontiguousarray(riverwall_elevation, dtype=np.float64)
        self.riverwall_rowIndex = np.ascontiguousarray(hydraulic_properties_rowIndex, dtype=np.int64)
        hp = np.ascontiguousarray(hydraulic_properties, dtype=np.float64)
        self.riverwall_hydraulic_properties = hp.reshape(-1)
        self.ncol_riverwall_hydraulic_properties = hp.shape[1] if hp.ndim == 2 else 5
        assert self.riverwall_elevation.size == idx.size == self.riverwall_rowIndex.size
        if self._dev is not None:
            self._release_device()

    def set_fractional_step_operator(self, operator):
        self.fractional_step_operators.append(operator)
        self._operators_dirty = True

    def print_operator_timestepping_statistics(self):
        """generic_domain.py:2320-2326"""
        for operator in self.fractional_step_operators:
            operator.print_timestepping_statistics()

    def print_operator_statistics(self):
        for operator in self.fractional_step_operators:
            operator.print_statistics()

    def get_centroid_coordinates(self, absolute=False):
        return self.centroid_coordinates

    def get_vertex_coordinates(self, absolute=False):
        return self.vertex_coordinates

    def get_edge_midpoint_coordinates(self, absolute=False):
        return self.edge_midpoint_coordinates

    def get_areas(self):
        return self.areas

    def get_time(self):
        return self.starttime + self.relative_time

    def set_time(self, time=0.0):
        """generic_domain.py:600-607: model time in absolute terms"""
        self.relative_time = float(time) - self.starttime
        if self._dev is not None:
            self._dev.set_time(self.relative_time)

    def set_relative_time(self, time=0.0):
        self.relative_time = float(time)
        if self._dev is not None:
            self._dev.set_time(self.relative_time)

    def get_relative_time(self):
        return self.relative_time

    def get_timestep(self):
        return self.timestep

    def get_boundary_flux_integral(self):
        return self.boundary_flux_integral

    def get_fractional_step_volume_integral(self):
        return self.fractional_step_volume_integral

    def compute_total_volume(self):
        self._pull_centroids()
        h = self.quantities["stage"].centroid_values - self.quantities["elevation"].centroid_values
        m = self.tri_full_flag == 1
        return float(np.sum(h[m] * self.areas[m]))

    get_water_volume = compute_total_volume

    # -- yield-time diagnostics on the host arrays (valid at yields; shallow_water_domain.py:1563-1625, 2611-2716)
    def get_wet_elements(self, indices=None, minimum_height=None):
        """indices (relative to `indices` if given) of the elements with depth > minimum_height (centroids)"""
        if minimum_height is None:
            minimum_height = _CONFIG["minimum_allowed_height"]
        self._pull_centroids()
        elevation = self.quantities["elevation"].get_values(location="centroids", indices=indices)
        stage = self.quantities["stage"].get_values(location="centroids", indices=indices)
        depth = stage - elevation
        return np.compress(depth > minimum_height, np.arange(len(depth)))

    def get_maximum_inundation_elevation(self, indices=None, minimum_height=None):
        """highest bed elevation among the wet elements"""
        wet = self.get_wet_elements(indices, minimum_height)
        return self.quantities["elevation"].get_maximum_value(indices=wet)

    def get_maximum_inundation_location(self, indices=None):
        wet = self.get_wet_elements(indices)
        return self.quantities["elevation"].get_maximum_location(indices=wet)

    def get_intersecting_segments(self, polyline, use_cache=False, verbose=False):
        """segments of a polyline (absolute coordinates) inside the triangles it crosses
        (neighbour_mesh.py:1124-1167)"""
        from .cross_section import Cross_section
        return Cross_section(self, polyline, verbose).segments

    def get_flow_through_cross_section(self, polyline, verbose=False):
        """total flow [m^3/s] across a polyline, left to right (shallow_water_domain.py:1750-1769)"""
        from .cross_section import Cross_section
        return Cross_section(self, polyline, verbose).get_flow_through_cross_section()

    def get_energy_through_cross_section(self, polyline, kind="total", verbose=False):
        """average energy head [m] along a polyline (shallow_water_domain.py:1772-1810)"""
        from .cross_section import Cross_section
        return Cross_section(self, polyline, verbose).get_energy_through_cross_section(kind)

    def compute_boundary_flows(self):
        """approximate flows across the boundary from the edge momenta (not the fluxes of evolve; see
        get_boundary_flux_integral for the exact figure): {tag: flow}, total inflow, total outflow"""
        uh = self.quantities["xmomentum"].get_values(location="edges")
        vh = self.quantities["ymomentum"].get_values(location="edges")
        flows, inflow, outflow = {}, 0.0, 0.0
        for (vol_id, edge_id), tag in self.boundary.items():
            momentum = [uh[vol_id, edge_id], vh[vol_id, edge_id]]
            normal = self.normals[vol_id, 2 * edge_id:2 * edge_id + 2]
            edge_flow = -(np.dot(momentum, normal) * self.edgelengths[vol_id, edge_id])
            if edge_flow > 0:
                inflow += edge_flow
            else:
                outflow += edge_flow
            flows[tag] = flows.get(tag, 0.0) + edge_flow
        return flows, inflow, outflow

    def volumetric_balance_statistics(self):
        flows, inflow, outflow = self.compute_boundary_flows()
        message = "---------------------------\nVolumetric balance report:\nNote: Boundary fluxes are not exact\n"
        message += "See get_boundary_flux_integral for exact computation\n--------------------------\n"
        message += "Total boundary inflow [m^3/s]: %.2f\n" % inflow
        message += "Total boundary outflow [m^3/s]: %.2f\n" % outflow
        message += "Net boundary flow by tags [m^3/s]\n"
        for tag in flows:
            message += "    %s [m^3/s]: %.2f\n" % (tag, flows[tag])
        message += "Total net boundary flow [m^3/s]: %.2f\n" % (inflow + outflow)
        message += "Total volume in domain [m^3]: %.2f\n" % self.compute_total_volume()
        return message

    def print_volumetric_balance_statistics(self):
        print(self.volumetric_balance_statistics())

    def report_water_volume_statistics(self, verbose=True, returnStats=False):
        """volume, boundary-flux integral, fractional-step volume integral and their balance
        (shallow_water_domain.py:2749-2781)"""
        vol = self.get_water_volume()
        if not hasattr(self, "_initial_volume"):
            self._initial_volume = vol
        bf, fs = self.get_boundary_flux_integral(), self.get_fractional_step_volume_integral()
        if verbose and self.processor == 0:
            print(" ")
            print("    Volume V is:", vol)
            print("    Boundary Flux integral BF: ", bf)
            print("    (rate + inlet) Fractional Step volume integral FS: ", fs)
            print("    V - BF - FS - InitialVolume :", vol - bf - fs - self._initial_volume)
            print(" ")
        if returnStats:
            return [vol, bf, fs]

    def statistics(self, *args, **kwargs):
        """mesh summary in the spirit of Mesh.statistics (neighbour_mesh.py:820-920)"""
        a = self.areas
        x, y = self.nodes[:, 0], self.nodes[:, 1]
        lines = ["------------------------------------------------",
                 "Mesh statistics:",
                 "  Number of triangles = %d" % self.number_of_triangles,
                 "  Extent [m]:",
                 "    x in [%e, %e]" % (x.min(), x.max()),
                 "    y in [%e, %e]" % (y.min(), y.max()),
                 "  Areas [m^2]:",
                 "    A in [%e, %e]" % (a.min(), a.max()),
                 "    number of distinct areas: %d" % len(np.unique(a)),
                 "  Boundary:",
                 "    Number of boundary segments == %d" % self.boundary_length,
                 "    Boundary tags == %s" % self.get_boundary_tags(),
                 "------------------------------------------------"]
        return "\n".join(lines)

    def print_statistics(self, *args, **kwargs):
        print(self.statistics())

    def get_nodes(self, absolute=False):
        return self.nodes

    def get_triangles(self, indices=None):
        return self.triangles if indices is None else self.triangles[np.asarray(indices, dtype=np.int64)]

    def get_number_of_triangles(self):
        return self.number_of_triangles

    # -- mesh accessors the reference forwards to its mesh (generic_domain.py:449-560, neighbour_mesh.py) ----
    def get_number_of_full_triangles(self, *args, **kwargs):
        return self.number_of_full_triangles

    def get_full_centroid_coordinates(self, absolute=False):
        return self.get_centroid_coordinates(absolute=absolute)[:self.number_of_full_triangles, :]

    def get_full_vertex_coordinates(self, absolute=False):
        return self.get_vertex_coordinates(absolute=absolute)[:3 * self.number_of_full_triangles, :]

    def get_full_triangles(self, *args, **kwargs):
        return self.get_triangles(*args, **kwargs)[:self.number_of_full_triangles, :]

    def get_full_nodes(self, absolute=False):
        return self.get_nodes(absolute=absolute)[:self.number_of_full_nodes, :]

    def get_disconnected_triangles(self):
        return np.reshape(np.arange(3 * self.number_of_triangles, dtype=int), (self.number_of_triangles, 3))

    def get_area(self):
        return np.sum(self.areas)

    def get_radii(self):
        return self.radii

    def get_extent(self, absolute=False):
        """xmin, xmax, ymin, ymax of the mesh"""
        C = self.get_vertex_coordinates(absolute=absolute)
        return np.min(C[:, 0]), np.max(C[:, 0]), np.min(C[:, 1]), np.max(C[:, 1])

    def get_vertex_coordinate(self, i, j, absolute=False):
        assert j in [0, 1, 2], "vertex id j must be an integer in [0,1,2]"
        return self.get_vertex_coordinates(absolute=absolute)[3 * i + j, :]

    def get_edge_midpoint_coordinate(self, i, j, absolute=False):
        assert j in [0, 1, 2], "edge midpoint id j must be an integer in [0,1,2]"
        return self.get_edge_midpoint_coordinates(absolute=absolute)[3 * i + j, :]

    def get_triangle_containing_point(self, point):
        """lowest id of a triangle that holds the point (neighbour_mesh.py:1057-1080)"""
        from .structures import triangle_containing_point
        return triangle_containing_point(self, point)

    def get_triangles_inside_polygon(self, polygon):
        """ids of the triangles whose centroid lies inside the polygon (neighbour_mesh.py:1083-1097)"""
        from .compat import inside_polygon
        return inside_polygon(self.get_centroid_coordinates(absolute=True), polygon)

    def get_evolved_quantities(self, vol_id, vertex=None, edge=None):
        """generic_domain.py:622-653; the evolved quantities of the shallow-water domain are its conserved ones"""
        if not (vertex is None or edge is None):
            raise Exception("Values for both vertex and edge was specified.Only one (or none) is allowed.")
        return self.get_conserved_quantities(vol_id, vertex=vertex, edge=edge)

    def get_starttime(self, datetime=False):
        return self.get_datetime(self.starttime) if datetime else self.starttime

    def get_evolve_starttime(self):
        return self.evolve_starttime

    def set_evolve_starttime(self, time):
        self.evolve_starttime = float(time)
        self.set_relative_time(self.evolve_starttime)

    def set_timezone(self, tz=None):
        """timezone of get_datetime: None (UTC), a tz database name or a ZoneInfo (shallow_water_domain.py:2453-2487)"""
        from zoneinfo import ZoneInfo
        if tz is None:
            self.timezone = ZoneInfo("UTC")
        elif isinstance(tz, str):
            self.timezone = ZoneInfo(tz)
        elif isinstance(tz, ZoneInfo):
            self.timezone = tz
        else:
            raise Exception("Unknown timezone %s" % tz)

    def get_timezone(self):
        if getattr(self, "timezone", None) is None:
            self.set_timezone()
        return self.timezone

    def get_datetime(self, timestamp=None):
        """the model time as a datetime in the domain's timezone (shallow_water_domain.py:2496-2516)"""
        from datetime import datetime, timezone
        if timestamp is None:
            timestamp = self.get_time()
        return datetime.fromtimestamp(timestamp, timezone.utc).astimezone(self.get_timezone())

    def set_institution(self, institution):
        self.institution = institution

    def evolve_to_end(self, finaltime=1.0):
        for _ in self.evolve(yieldstep=None, finaltime=finaltime):
            pass

    def write_time(self, track_speeds=False):
        print(self.timestepping_statistics(track_speeds))

    def get_number_of_nodes(self):
        return self.number_of_nodes

    def get_normal(self, i, j):
        return self.normals[i, 2 * j:2 * j + 2]

    def get_conserved_quantities(self, vol_id, vertex=None, edge=None):
        """stage, xmomentum, ymomentum of one triangle at its centroid, a vertex or an edge
        (generic_domain.py:585-620)"""
        assert vertex is None or edge is None, "Values for both vertex and edge was specified."
        if vertex is None and edge is None:
            self._pull_centroids()
        out = np.zeros(3)
        for k, name in enumerate(self.conserved_quantities):
            q = self.quantities[name]
            if vertex is not None:
                out[k] = q.vertex_values[vol_id, vertex]
            elif edge is not None:
                out[k] = q.edge_values[vol_id, edge]
            else:
                out[k] = q.centroid_values[vol_id]
        return out

    def timestepping_statistics(self, *a, **k):
        msg = "Time = %.4f (sec), " % self.get_time()
        if self.recorded_min_timestep == self.recorded_max_timestep:
            msg += "delta t = %.8f (s), " % self.recorded_min_timestep
        elif self.recorded_min_timestep > self.recorded_max_timestep:
            msg += "delta t = %.8f (s), " % self.recorded_min_timestep
        else:
            msg += "delta t in [%.8f, %.8f] (s), " % (self.recorded_min_timestep, self.recorded_max_timestep)
        msg += "steps=%d" % self.number_of_steps
        return msg

    def print_timestepping_statistics(self, *a, **k):
        print(self.timestepping_statistics(*a, **k))

    # ------------------------------------------------------------------
    # device plumbing
    # ------------------------------------------------------------------
    def _param_dict(self):
        return dict(
            epsilon=self.epsilon, H0=self.H0, g=self.g, minimum_allowed_height=self.minimum_allowed_height,
            maximum_allowed_speed=self.maximum_allowed_speed, evolve_max_timestep=self.evolve_max_timestep,
            evolve_min_timestep=self.evolve_min_timestep, beta_w=self.beta_w, beta_w_dry=self.beta_w_dry,
            beta_uh=self.beta_uh, beta_uh_dry=self.beta_uh_dry, beta_vh=self.beta_vh,
            beta_vh_dry=self.beta_vh_dry, CFL=self.CFL, fixed_flux_timestep=self.fixed_flux_timestep,
            extrapolate_velocity_second_order=self.extrapolate_velocity_second_order,
            low_froude=self.low_froude, timestepping_method=self.timestepping_method,
            use_sloped_mannings=self.use_sloped_mannings, max_smallsteps=self.max_smallsteps,
            default_order=self.default_order, ghost_layer_width=self.ghost_layer_width,
            centroid_transmissive_bc=self.centroid_transmissive_bc,
            track_max_speed=getattr(self, "track_max_speed", False),
        )

    def _mesh_dict(self):
        m = self.mesh
        d = dict(
            neighbours=m.neighbours, neighbour_edges=m.neighbour_edges,
            surrogate_neighbours=m.surrogate_neighbours, number_of_boundaries=m.number_of_boundaries,
            tri_full_flag=self.tri_full_flag, normals=m.normals, edgelengths=m.edgelengths,
            radii=m.radii, areas=m.areas, centroid_coordinates=m.centroid_coordinates,
            edge_coordinates=m.edge_midpoint_coordinates, vertex_coordinates=m.vertex_coordinates,
            boundary_cells=m.boundary_cells, boundary_edges=m.boundary_edges,
        )
        for k in ("edge_flux_type", "edge_river_wall_counter", "riverwall_elevation", "riverwall_rowIndex",
                  "riverwall_hydraulic_properties", "number_of_riverwall_edges",
                  "ncol_riverwall_hydraulic_properties"):
            if hasattr(self, k):
                d[k] = getattr(self, k)
        return d

    def _release_device(self):
        if self._dev is not None:
            self._pull_centroids()
            self._dev.close()
            self._dev = None

    def _ensure_device(self):
        if self._dev is None:
            perm = None
            if self.reorder:
                perm = self._locality_permutation()
            self._dev = _b.DeviceDomain(self._mesh_dict(), self._param_dict(), device=self.device,
                                        permutation=perm)
            self._params_dirty = False
            self._boundary_dirty = True
            self._operators_dirty = True
            self._segments = {}
            for q in self.quantities.values():
                q.host_dirty = True
            for op in self.fractional_step_operators:
                if not getattr(op, "host_side", False):
                    op.op_id = None
            if self.processor in self.full_send_dict and self.processor in self.ghost_recv_dict:
                self._dev.set_local_ghost_copy(self.full_send_dict[self.processor][0],
                                               self.ghost_recv_dict[self.processor][0])
            if getattr(self, "_comm", None) is not None:
                self._attach_device_comm()
        if self._params_dirty:
            self._dev.set_params(self._param_dict())
            self._params_dirty = False
        return self._dev

    def attach_communicator(self, comm):
        """Join a process group (anuga_core_b200.parallel.Communicator): the halo lists of
        full_send_dict / ghost_recv_dict and the global timestep then run over NCCL inside the
        device time loop (Parallel_domain.update_ghosts / update_timestep,
        parallel_shallow_water.py:128-144)."""
        self._comm = comm
        if self._dev is not None:
            self._attach_device_comm()

    def _attach_device_comm(self):
        import os
        comm = self._comm
        if comm.size <= 1:
            return
        from .parallel import nccl_library_path
        lib = nccl_library_path()
        if lib and "SWK_NCCL_LIB" not in os.environ:
            os.environ["SWK_NCCL_LIB"] = lib
        if getattr(comm, "nccl", None) is not None:      # the process-level NCCL communicator
            self._dev.comm_attach(comm.nccl)
        else:                                            # a foreign process group ships the id
            uid = _b.DeviceDomain.nccl_unique_id() if comm.rank == 0 else b""
            uid = comm.broadcast_bytes(uid, 128)
            self._dev.comm_init(uid, comm.rank, comm.size)
        send = {int(q): v[0] for q, v in self.full_send_dict.items() if q != self.processor}
        recv = {int(q): v[0] for q, v in self.ghost_recv_dict.items() if q != self.processor}
        self._dev.set_halo(send, recv)

    def _locality_permutation(self):
        """Morton order of the centroids; full triangles stay in front of ghosts so that
        the reference's "full first, ghosts after" numbering invariant survives."""
        perm = morton_order(self.centroid_coordinates)
        full = self.tri_full_flag[perm] == 1
        # halo-source triangles (listed in any full_send_dict) lead the numbering, so that the device
        # can update them first and ship them while the interior is still being computed
        send = np.zeros(self.number_of_triangles, dtype=bool)
        for q, v in self.full_send_dict.items():
            if q != self.processor:
                send[np.asarray(v[0], dtype=np.int64)] = True
        s = send[perm]
        return np.concatenate([perm[full & s], perm[full & ~s], perm[~full]])

    _DEVICE_NAMES = {"stage": "STAGE", "xmomentum": "XMOM", "ymomentum": "YMOM",
                     "elevation": "ELEVATION", "height": "HEIGHT"}

    def _push_quantities(self, force=False):
        dev = self._ensure_device()
        q = self.quantities
        for name, qid in (("stage", "STAGE_C"), ("xmomentum", "XMOM_C"), ("ymomentum", "YMOM_C"),
                          ("elevation", "ELEVATION_C"), ("friction", "FRICTION_C")):
            if force or q[name].host_dirty:
                if self.pin_host_arrays and name in self.conserved_quantities:
                    dev.pin(q[name].centroid_values)       # moved every yield: page-lock once
                dev.set_quantity(qid, q[name].centroid_values)
                q[name].host_dirty = False
                self._stale.discard(name)

    def sync_from_host(self, quantities=("stage", "xmomentum", "ymomentum", "elevation", "friction")):
        """Declare that the numpy centroid arrays were modified in place."""
        for name in quantities:
            self.quantities[name].host_dirty = True
        self._push_quantities()

    def _pull_centroids(self):
        if self._dev is None:
            return
        q = self.quantities
        for name, qid in (("stage", "STAGE_C"), ("xmomentum", "XMOM_C"), ("ymomentum", "YMOM_C")):
            if name in self._stale and not q[name].host_dirty:
                self._dev.get_quantity(qid, out=q[name].centroid_values)
                self._stale.discard(name)

    def sync_to_host(self):
        self._pull_centroids()

    def _mark_device_newer(self):
        self._stale = {"stage", "xmomentum", "ymomentum"}
        self._lazy_stale = {(n, a) for n in ("stage", "xmomentum", "ymomentum", "elevation", "height")
                            for a in ("edge_values", "vertex_values")}
        self._lazy_stale |= {(n, "explicit_update") for n in ("stage", "xmomentum", "ymomentum")}

    def _fetch_lazy(self, quantity, array_name, out):
        key = (quantity.name, array_name)
        stale = getattr(self, "_lazy_stale", None)
        if self._dev is None or not stale or key not in stale:
            return
        base = self._DEVICE_NAMES.get(quantity.name)
        suffix = {"edge_values": "_E", "vertex_values": "_V", "explicit_update": "_EU"}[array_name]
        self._dev.get_quantity(base + suffix, out=out)
        stale.discard(key)

    def _push_boundaries(self, t):
        dev = self._dev
        if self.boundary_map is None:
            raise Exception("Boundary tags must be bound to boundary objects before evolving system, "
                            "e.g. using the method set_boundary.\nThis system has the boundary tags %s"
                            % self.get_boundary_tags())
        if self._boundary_dirty:
            self._segments = {}
            for seg, tag in enumerate(sorted(self.tag_boundary_cells.keys())):
                B = self.boundary_map[tag]
                ids = self.tag_boundary_cells[tag]
                kind = _b.BC_NONE if B is None else B.device_kind
                if hasattr(B, "frames_for"):          # time-space table kinds: frames resident on the device
                    dev.set_boundary_segment(seg, kind, ids, (0.0, 0.0, 0.0))
                    dev.set_boundary_table(seg, B.frames_for(ids))
                    dev.set_boundary_values(seg, self._segment_values(seg, tag, B, 0, t))
                else:
                    vals = (0.0, 0.0, 0.0) if B is None else B.device_values(t)
                    dev.set_boundary_segment(seg, kind, ids, vals)
                self._segments[tag] = (seg, B)
            self._boundary_dirty = False
        else:
            for tag, (seg, B) in list(self._segments.items()):
                if B is not None and B.time_dependent:
                    dev.set_boundary_values(seg, self._segment_values(seg, tag, B, 0, t))

    def _segment_values(self, seg, tag, B, substep, t):
        """value-table entries of a time-dependent boundary at time t; when its data has run out
        (Modeltime_too_late) and it carries a default_boundary OBJECT, that boundary takes the segment over
        (generic_boundary_conditions.py:486-516, 655-684)"""
        try:
            return B.values_for_substep(self._dev, seg, substep, t, self.tag_boundary_cells[tag])
        except BaseException as e:
            default = getattr(B, "default_boundary", None)
            if type(e).__name__ != "Modeltime_too_late" or not hasattr(default, "device_kind"):
                raise
            if getattr(B, "verbose", False) and not getattr(B, "default_boundary_invoked", False):
                print("%s\nInstead I will use the default boundary: %s\nNote: Further warnings will be supressed"
                      % (e, default))
            B.default_boundary_invoked = True
            self._dev.set_boundary_segment(seg, default.device_kind, self.tag_boundary_cells[tag],
                                           default.device_values(t))
            self._segments[tag] = (seg, default)
            return default.device_values(t)

    def _push_operators(self, t):
        dev = self._dev
        for op in self.fractional_step_operators:
            if getattr(op, "host_side", False):
                continue
            if op.op_id is None:
                op.op_id = dev.add_rate_operator(op.current_rate(t), op.current_factor(t),
                                                 op.rate_array, op.indices)
                if op.time_dependent:       # scalars read from the device value table, refreshed every step
                    dev.set_rate_dynamic(op.op_id, True)
            elif op.time_dependent:
                dev.set_rate(op.op_id, op.current_rate(t), op.current_factor(t))
        self._operators_dirty = False

    def _device_forcing_terms(self):
        out = []
        for f in self.forcing_terms:
            if hasattr(f, "explicit_forcing"):
                out.append(f)
            elif getattr(f, "__name__", "") not in ("manning_friction_implicit", "manning_friction_explicit"):
                raise NotImplementedError("forcing term %r has no device implementation (SURVEY.md 8(f))" % (f,))
        return out

    def _push_forcing(self, t, force=False):
        """compute_forcing_terms (generic_domain.py:2417-2427) for the state-independent terms: their
        per-triangle momentum forcing is evaluated on the host and handed to the update kernels"""
        terms = self._device_forcing_terms()
        if not terms:
            if self._forcing_on_device:
                self._dev.set_explicit_forcing(None, None, None)
                self._forcing_on_device = False
            return
        if self._forcing_on_device and not force and not any(f.time_dependent for f in terms):
            return
        # the terms add to the explicit updates one after the other, in list order (generic_domain.py:2426)
        F = {name: np.zeros(self.number_of_triangles) for name in self.conserved_quantities}
        for f in terms:
            f.explicit_forcing(self, t, F)
        self._dev.set_explicit_forcing(F["stage"], F["xmomentum"], F["ymomentum"])
        self._forcing_on_device = True

    def _forcing_depends_on_time(self):
        return any(f.time_dependent for f in self._device_forcing_terms())

    def _needs_host_stepping(self):
        if self._forcing_depends_on_time():
            return True
        if self.boundary_map:
            for B in self.boundary_map.values():
                if B is not None and B.time_dependent:
                    return True
        return any(op.time_dependent for op in self.fractional_step_operators)

    # ------------------------------------------------------------------
    # individual passes on resident data (same names as the reference's methods)
    # ------------------------------------------------------------------
    def distribute_to_vertices_and_edges(self):
        self._push_quantities()
        self._dev.distribute_to_vertices_and_edges()
        self._mark_device_newer()

    def protect_against_infinitesimal_and_negative_heights(self):
        self._push_quantities()
        me = self._dev.protect()
        self._mark_device_newer()
        return me

    def update_boundary(self):
        self._ensure_device()
        self._push_boundaries(self.get_time())
        self._dev.update_boundary()

    def compute_fluxes(self, substep=0):
        self._ensure_device()
        self.flux_timestep = self._dev.compute_fluxes(substep)
        self._lazy_stale = getattr(self, "_lazy_stale", set()) | \
            {(n, "explicit_update") for n in ("stage", "xmomentum", "ymomentum")}
        return self.flux_timestep

    def update_conserved_quantities(self):
        n = self._dev.update_conserved_quantities(self.timestep)
        self._mark_device_newer()
        if n > 0:
            import warnings
            warnings.warn("Negative cells being set to zero depth, possible loss of conservation. \n"
                          "Consider using domain.report_water_volume_statistics() to check the extent "
                          "of the problem")

    def backup_conserved_quantities(self):
        self._push_quantities()
        self._dev.backup_conserved_quantities()

    def saxpy_conserved_quantities(self, a, b, divide_by=1.0):
        self._dev.saxpy_conserved_quantities(a, b, divide_by)
        self._mark_device_newer()

    def update_ghosts(self):
        self._ensure_device()
        self._dev.update_ghosts()
        self._mark_device_newer()

    def set_track_max_speed(self, flag=True):
        """max_speed[k] (sw_domain_openmp.c:709-710) is a diagnostic; the device time loop only
        writes it when asked (8 bytes per triangle and step)."""
        self.track_max_speed = bool(flag)
        self._params_dirty = True

    def get_max_speed(self):
        self._ensure_device()
        return self._dev.get_quantity("MAX_SPEED")

    @property
    def max_speed(self):
        return self.get_max_speed()

    # ------------------------------------------------------------------
    # evolve
    # ------------------------------------------------------------------
    def update_timestep(self, yieldstep, finaltime):
        """generic_domain.py:2349-2415 (host version, used by the host-stepped path)"""
        if self.fixed_flux_timestep is not None:
            self.flux_timestep = self.fixed_flux_timestep
        timestep = min(self.CFL * self.flux_timestep, self.evolve_max_timestep)
        self.recorded_max_timestep = max(timestep, self.recorded_max_timestep)
        self.recorded_min_timestep = min(timestep, self.recorded_min_timestep)
        if timestep < self.evolve_min_timestep:
            self.smallsteps += 1
            if self.smallsteps > self.max_smallsteps:
                self.smallsteps = 0
                if self._order_ == 1:
                    raise Exception("WARNING: Too small timestep %.16f reached even after %d steps of 1 order scheme"
                                    % (timestep, self.max_smallsteps))
                self._order_ = 1
        else:
            self.smallsteps = 0
            if self._order_ == 1 and self.default_order == 2:
                self._order_ = 2
        if self.relative_finaltime is not None and self.relative_time + timestep > self.relative_finaltime:
            timestep = self.relative_finaltime - self.relative_time
        if self.relative_time + timestep > self.relative_yieldtime:
            timestep = self.relative_yieldtime - self.relative_time
        self.timestep = timestep

    def _host_step(self):
        """One timestep driven from Python (evolve_one_*_step, generic_domain.py:1914-2179):
        used when boundary or operator callbacks must be evaluated on the host between
        substeps.  Data stays resident; only scalars cross the bus."""
        dev = self._dev
        method = self.timestepping_method
        t0 = self.relative_time

        def substep(k):
            dev.distribute_to_vertices_and_edges()
            self._push_boundaries(self.get_time())
            dev.update_boundary()
            self._push_forcing(self.get_time())        # compute_forcing_terms at the substep's time
            return dev.compute_fluxes(k)

        if method != "euler":
            dev.backup_conserved_quantities()
        self.flux_timestep = substep(0)
        self.update_timestep(self.yieldstep, self.finaltime)
        dt = self.timestep
        self._check_negative(dev.update_conserved_quantities(dt))
        if method == "rk2":
            self.relative_time = t0 + dt
            if self.ghost_layer_width < 4:
                dev.update_ghosts()
            substep(1)
            self._check_negative(dev.update_conserved_quantities(dt))
            dev.saxpy_conserved_quantities(0.5, 0.5)
        elif method == "rk3":
            self.relative_time = t0 + dt
            dev.update_ghosts()
            substep(1)
            self._check_negative(dev.update_conserved_quantities(dt))
            dev.saxpy_conserved_quantities(0.25, 0.75)
            self.relative_time = t0 + dt * 0.5
            dev.update_ghosts()
            substep(2)
            self._check_negative(dev.update_conserved_quantities(dt))
            dev.saxpy_conserved_quantities(2.0, 1.0, 3.0)
            self.relative_time = t0 + dt
        # NB the reference leaves relative_time advanced by dt after an rk2 / rk3 step (set_relative_time,
        # generic_domain.py:2011, 2176) and only _evolve_base resets it after apply_fractional_steps
        # (:1849-1855): fractional-step operators therefore see t0 + dt (euler: t0).  Kept.

    def _host_step_with_operators(self):
        """evolve_one_*_step + apply_fractional_steps; time-dependent rates are evaluated at the
        time the reference's operators see (see _host_step)"""
        self._host_step()
        self._push_operators(self.get_time())
        self._host_fractional_steps()

    def _check_negative(self, n):
        self._negative_cells = getattr(self, "_negative_cells", 0) + n

    def evolve(self, yieldstep=None, outputstep=None, finaltime=None, duration=None, skip_initial_step=False):
        """Domain.evolve (shallow_water_domain.py:2300-2407): the time loop below plus, with
        set_store(True), the SWW file: created before the first yield, one frame every `outputstep`."""
        if outputstep is None or yieldstep is None:
            self.output_frequency = 1
        else:
            freq = outputstep / yieldstep
            assert float(freq).is_integer(), \
                "outputstep (%s) should be an integer multiple of yieldstep (%s)" % (outputstep, yieldstep)
            self.output_frequency = int(freq)
        new_file = self.store and (self.relative_time == 0.0 or not self.evolved_called)
        if new_file and (skip_initial_step or self.evolved_called):
            # no initial yield to hang the file creation on: extrapolate once now (as the reference does)
            dev = self._ensure_device()
            self._push_quantities(force=not hasattr(self, "_pushed_once"))
            self._pushed_once = True
            dev.distribute_to_vertices_and_edges()
            self._mark_device_newer()
            self._pull_centroids()
            self.initialise_storage()
            new_file = False
        for t in self._evolve_base(yieldstep=yieldstep, finaltime=finaltime, duration=duration,
                                   skip_initial_step=skip_initial_step):
            if self.store:
                if new_file:
                    self.initialise_storage()
                    new_file = False
                if self.yieldstep_counter % self.output_frequency == 0:
                    self.store_timestep()
            if self.checkpoint:
                self._checkpoint_if_due()
            yield t
            self.yieldstep_counter += 1

    def _evolve_base(self, yieldstep=None, finaltime=None, duration=None, skip_initial_step=False):
        """Generator with the reference's protocol: yields the model time with the
        conserved centroid arrays valid on the host (generic_domain.py:1715-1912)."""
        if self.boundary_map is None:
            raise Exception("Boundary tags must be bound to boundary objects before evolving system, "
                            "e.g. using the method set_boundary.\nThis system has the boundary tags %s"
                            % self.get_boundary_tags())
        if self.evolved_called:
            skip_initial_step = True
        elif not hasattr(self, "_initial_volume"):
            self._initial_volume = self.compute_total_volume()      # volume_history[0] of the reference
        self.evolved_called = True
        if skip_initial_step:
            self.evolve_starttime = self.relative_time
        if self.relative_time != self.evolve_starttime:
            self.relative_time = self.evolve_starttime
        yieldstep = self.evolve_max_timestep if yieldstep is None else float(yieldstep)
        self.yieldstep = yieldstep
        self._order_ = self.default_order
        if finaltime is not None and duration is not None:
            raise Exception("Only one of finaltime and duration may be specified")
        if finaltime is not None:
            self.finaltime = float(finaltime)
            self.relative_finaltime = self.finaltime - self.starttime
        if duration is not None:
            self.finaltime = float(duration) + self.get_time()
            self.relative_finaltime = float(duration) + self.relative_time
        if self.relative_finaltime is not None and self.relative_finaltime < self.relative_time:
            import warnings
            warnings.warn("\n finaltime %g is less than current time %g! finaltime set to current time"
                          % (self.finaltime, self.get_time()))
            self.finaltime = self.get_time()
            self.relative_finaltime = self.relative_time
            return

        if self.numproc > 1 and getattr(self, "_comm", None) is None and \
                any(int(p) != self.processor for p in list(self.full_send_dict) + list(self.ghost_recv_dict)):
            raise Exception("this sub-domain exchanges halos with other ranks but has no communicator: call "
                            "domain.attach_communicator(parallel.communicator()) (e.g. after unpickling it)")
        dev = self._ensure_device()
        self._push_quantities(force=not hasattr(self, "_pushed_once"))
        self._pushed_once = True
        dev.set_time(self.relative_time)
        self._push_boundaries(self.get_time())
        self._push_operators(self.get_time())
        self._push_forcing(self.get_time(), force=True)

        self.relative_yieldtime = self.relative_time + yieldstep
        self.recorded_min_timestep = self.evolve_max_timestep
        self.recorded_max_timestep = self.evolve_min_timestep
        self.number_of_steps = 0
        self.number_of_first_order_steps = 0
        dev.reset_yield_statistics()
        dev.update_ghosts()

        if not skip_initial_step:
            dev.distribute_to_vertices_and_edges()
            dev.update_boundary()
            self._mark_device_newer()
            self._pull_centroids()
            yield self.get_time()

        while True:
            # user code may have touched the arrays during the yield
            self._push_quantities()
            if self._boundary_dirty:
                self._push_boundaries(self.get_time())
            if self._operators_dirty:
                self._push_operators(self.get_time())
            if self._params_dirty:
                dev.set_params(self._param_dict())
                self._params_dirty = False

            path = self._evolve_path()
            if path == 1:
                reason = self._evolve_device_steps_host_operators()
            elif path == 2:
                reason = self._evolve_split_steps()
            elif path == 3:
                reason = self._evolve_host_stepped()
            else:
                r = dev.evolve(self.relative_yieldtime, self.relative_finaltime, 0)
                self._absorb(r)
                reason = r.stop_reason
            self._mark_device_newer()
            self._pull_centroids()
            if getattr(self, "_negative_cells_warned", 0) < self._negative_total():
                import warnings
                self._negative_cells_warned = self._negative_total()
                warnings.warn("Negative cells being set to zero depth, possible loss of conservation. \n"
                              "Consider using domain.report_water_volume_statistics() to check the extent "
                              "of the problem")
            if reason == 2:
                yield self.get_time()
                break
            yield self.get_time()
            self.relative_yieldtime += yieldstep
            self.recorded_min_timestep = self.evolve_max_timestep
            self.recorded_max_timestep = self.evolve_min_timestep
            self.number_of_steps = 0
            self.number_of_first_order_steps = 0
            dev.reset_yield_statistics()

    def _negative_total(self):
        return getattr(self, "_negative_device", 0) + getattr(self, "_negative_cells", 0)

    def _absorb(self, r):
        self.relative_time = r.time
        self.timestep = r.timestep
        self.flux_timestep = r.flux_timestep
        self.recorded_min_timestep = r.recorded_min_timestep
        self.recorded_max_timestep = r.recorded_max_timestep
        self.number_of_steps = r.number_of_steps
        self.number_of_first_order_steps = r.number_of_first_order_steps
        self.boundary_flux_integral = r.boundary_flux_integral
        self.fractional_step_volume_integral = r.fractional_step_volume_integral
        self.total_steps = r.total_steps
        self.mass_error = r.mass_error
        self._negative_device = r.negative_cells
        self.kernel_launches += r.kernel_launches

    def _evolve_path(self):
        """Which time loop runs this yield segment: 0 device-resident, 1 device steps + host-side operators,
        2 two-half steps (time-dependent boundary values / rates), 3 host-stepped passes.  Under a
        communicator the choice is COLLECTIVE: the loops differ in how many steps they launch ahead, so every
        rank must run the same one even if only some ranks hold the time-dependent boundary or the inlet."""
        if not self._needs_host_stepping():
            mine = 0
        elif self._only_host_side_operators_need_the_host():
            mine = 1
        elif self._only_values_depend_on_time():
            mine = 2
        else:
            mine = 3
        comm = getattr(self, "_comm", None)
        if comm is None or comm.size <= 1:
            return mine
        # flags: bit 0 = somebody needs path 1, bit 1 = path 2, bit 2 = path 3
        seen = 0
        for bit, p in enumerate((1, 2, 3)):
            if comm.allreduce_max(1.0 if mine == p else 0.0) > 0.0:
                seen |= 1 << bit
        if seen == 0:
            return 0
        if seen in (1, 2):
            return 1 if seen == 1 else 2
        return 3

    def _only_host_side_operators_need_the_host(self):
        """static boundaries and rates: the step itself can stay in the device time loop and only the
        host-side operators (inlets, culverts) run between steps"""
        if self._forcing_depends_on_time():
            return False
        if self.boundary_map:
            for B in self.boundary_map.values():
                if B is not None and B.time_dependent:
                    return False
        return all(getattr(op, "host_side", False) or not op.time_dependent
                   for op in self.fractional_step_operators)

    def _only_values_depend_on_time(self):
        """time-dependent boundary values / rate(t) of device operators, no host-side operator: the fused
        device step runs in two halves with one host visit per step (_evolve_split_steps)"""
        return not self._forcing_depends_on_time() and \
            not any(getattr(op, "host_side", False) for op in self.fractional_step_operators)

    def _evolve_split_steps(self):
        """_evolve_base's while loop for boundary values / rates that are Python functions of time
        (boundaries.py:384-517, 553-635; generic_boundary_conditions.py:297-411; rate_operators.py:276-289).
        The device runs the same fused, graph-replayed step as the resident loop, in two halves
        (include/swk.h: swk_step_*): after the first half the timestep exists, the host evaluates its
        functions at the times the reference's substeps see - t + dt (rk2, rk3), t + dt/2 (rk3's third
        substep), and t + dt for the next step's first substep (generic_domain.py:2011, 2093, 2132) -
        and the second half is released.  One D2H scalar read and one small H2D copy per step."""
        dev = self._dev
        method = self.timestepping_method
        segs = [(seg, tag) for tag, (seg, B) in self._segments.items() if B is not None and B.time_dependent]

        def values(seg, tag, substep, t):
            B = self._segments[tag][1]
            if not B.time_dependent:        # a default boundary took over
                return B.device_values(t)
            return self._segment_values(seg, tag, B, substep, t)
        ops = [op for op in self.fractional_step_operators
               if not getattr(op, "host_side", False) and op.time_dependent]
        dev.step_begin(self.relative_yieldtime, self.relative_finaltime)
        steps = 0
        while True:
            r = dev.step_first()
            if r.stop_reason != 0:
                break
            t0, dt = r.time, r.timestep
            steps += 1
            if self.record_timestep_history:
                self.timestep_history.append(dt)
            self.timestep = dt
            self.relative_time = t0 + dt
            T1 = self.get_time()
            if method != "euler":
                for seg, tag in segs:
                    dev.set_boundary_values_substep(seg, 1, values(seg, tag, 1, T1))
            if method == "rk3":
                self.relative_time = t0 + dt * 0.5
                T2 = self.get_time()
                for seg, tag in segs:
                    dev.set_boundary_values_substep(seg, 2, values(seg, tag, 2, T2))
            # fractional-step operators see t0 after an euler step and t0 + dt after rk2 / rk3 (see _host_step)
            self.relative_time = t0 if method == "euler" else t0 + dt
            Top = self.get_time()
            for op in ops:
                dev.set_rate(op.op_id, op.current_rate(Top), op.current_factor(Top))
            # the next step's first substep (and the yield's update_boundary) see the time after this step
            self.relative_time = t0 + dt
            for seg, tag in segs:
                dev.set_boundary_values_substep(seg, 0, values(seg, tag, 0, T1))
            dev.step_rest()
        reason = r.stop_reason
        if reason == 2:
            # time snapped to finaltime (generic_domain.py:1870-1888): the yield's boundary update sees it
            self.relative_time = r.time
            for seg, tag in segs:
                dev.set_boundary_values(seg, values(seg, tag, 0, self.get_time()))
        r = dev.step_end()
        self._absorb(r)
        return reason

    def _evolve_device_steps_host_operators(self):
        """_evolve_base's while loop when only host-side operators (inlets, culverts, rate(x, y, t)) need the
        host: per timestep one fused device step - launched in its two halves, the host learning (t, dt) in
        between with the step's only synchronisation - followed by the host-side operators on their registered
        cell sets (one small gather, the scalar hydraulics, a queued scatter) and the ghost update."""
        dev = self._dev
        host_ops = [op for op in self.fractional_step_operators if getattr(op, "host_side", False)]
        dev.step_begin(self.relative_yieldtime, self.relative_finaltime)
        while True:
            r = dev.step_first()
            if r.stop_reason != 0:
                break
            dev.step_rest()
            self._host_operators_after_step(host_ops, r.time, r.timestep)
        reason = r.stop_reason
        r = dev.step_end()          # the yield's extrapolation sees the state the operators left
        self._absorb(r)
        return reason

    def _host_operators_after_step(self, host_ops, t0, dt):
        dev = self._dev
        self.timestep = dt
        if self.record_timestep_history:
            self.timestep_history.append(dt)
        # operators see t0 after an euler step and t0 + dt after rk2 / rk3 (see _host_step)
        self.relative_time = t0 if self.timestepping_method == "euler" else t0 + dt
        for op in host_ops:
            added = op()
            if added != 0.0:
                dev.add_fractional_step_volume(added)
        self.relative_time = t0 + dt
        dev.update_ghosts_async()

    def run_steps_with_host_operators(self, n_steps):
        """Exactly n_steps timesteps of the loop above (benchmarks): returns the elapsed milliseconds between
        two device synchronisations (the host-side hydraulics are part of the step, so wall clock)."""
        import time as _time
        dev = self._dev
        host_ops = [op for op in self.fractional_step_operators if getattr(op, "host_side", False)]
        dev.synchronize()
        t0 = _time.perf_counter()
        dev.step_begin(1.0e300, None)
        for _ in range(int(n_steps)):
            r = dev.step_first()
            dev.step_rest()
            self._host_operators_after_step(host_ops, r.time, r.timestep)
        dev.synchronize()
        ms = (_time.perf_counter() - t0) * 1.0e3
        self._absorb(dev.step_end())
        self._mark_device_newer()
        return ms

    def _evolve_host_stepped(self):
        """_evolve_base's while loop with host-evaluated callbacks (time-dependent
        boundaries, rate functions of time).  One device->host scalar read per step."""
        dev = self._dev
        while True:
            t0 = self.relative_time
            dev.set_time(t0)
            self._push_operators(self.get_time())
            # the whole step on the device when nothing depends on the substep time,
            # else substep by substep
            self._host_step_with_operators()
            self.relative_time = t0 + self.timestep
            dev.set_time(self.relative_time)
            dev.update_ghosts()
            self.number_of_steps += 1
            self.total_steps += 1
            if self.record_timestep_history:
                self.timestep_history.append(self.timestep)
            if self._order_ == 1:
                self.number_of_first_order_steps += 1
            if self.relative_finaltime is not None and self.relative_time >= self.relative_finaltime - self.epsilon:
                if self.relative_time > self.relative_finaltime:
                    raise Exception("WARNING (domain.py): time overshot finaltime. ")
                self.relative_time = self.relative_finaltime
                dev.set_time(self.relative_time)
                dev.distribute_to_vertices_and_edges()
                self._push_boundaries(self.get_time())
                dev.update_boundary()
                return 2
            if self.relative_time >= self.relative_yieldtime:
                dev.distribute_to_vertices_and_edges()
                self._push_boundaries(self.get_time())
                dev.update_boundary()
                return 1

    def _host_fractional_steps(self):
        """apply_fractional_steps (generic_domain.py:2312) for a host-driven step: the device
        operators (boundary-flux integral, Rate_operators) first, then host-side operators in
        registration order on gathered cells.  (Mixed orders are not supported.)"""
        self._dev.apply_fractional_steps(self.timestep)
        for op in self.fractional_step_operators:
            if getattr(op, "host_side", False):
                added = op()
                if added != 0.0:
                    self._dev.add_fractional_step_volume(added)
        r = self._dev.get_statistics()
        self.boundary_flux_integral = r.boundary_flux_integral
        self.fractional_step_volume_integral = r.fractional_step_volume_integral
        self.mass_error = r.mass_error


def load_checkpoint_file(domain_name="domain", checkpoint_dir=".", time=None):
    """the most recent (or the given) checkpoint of a run; with several ranks every rank loads its own file
    and all fall back to an earlier time together if one of them cannot (shallow_water/checkpoint.py:24-76)"""
    import glob
    import os
    from . import parallel
    try:
        import dill as pickle
    except ImportError:
        import pickle
    if parallel.numprocs > 1:
        domain_name = domain_name + "_P{}_{}".format(parallel.numprocs, parallel.myid)
    if time is None:
        times = set()
        for path in glob.glob(os.path.join(checkpoint_dir, domain_name) + "_*.pickle"):
            try:
                times.add(float(os.path.basename(path)[len(domain_name) + 1:-len(".pickle")]))
            except ValueError:
                pass
        times = sorted(times)
    else:
        times = [float(time)]
    if len(times) == 0:
        raise Exception("Unable to open checkpoint file")
    for t in reversed(times):
        name = os.path.join(checkpoint_dir, domain_name) + "_" + str(t) + ".pickle"
        try:
            with open(name, "rb") as f:
                domain = pickle.load(f)
            ok = True
        except Exception:
            domain, ok = None, False
        if parallel.numprocs > 1:
            ok = parallel.communicator().allreduce_max(0.0 if ok else 1.0) == 0.0
        if ok:
            if parallel.numprocs > 1:
                # a restored Parallel_domain keeps communicating in the reference (global MPI layer); here
                # the process group is not picklable, so the restored sub-domain joins it again
                domain.attach_communicator(parallel.communicator())
            return domain
    raise Exception("Unable to open checkpoint file")


def rectangular_cross_domain(m, n, len1=1.0, len2=1.0, origin=(0.0, 0.0), **kwargs):
    """anuga/extras.py:13 - rectangular_cross mesh wrapped in a Domain."""
    points, vertices, boundary = rectangular_cross(int(m), int(n), len1, len2, origin)
    mesh = Mesh(points, vertices, boundary, use_inscribed_circle=kwargs.pop("use_inscribed_circle", False),
                neighbour_structure=rectangular_cross_neighbours(int(m), int(n)))
    return Domain(mesh=mesh, **kwargs)
