"""Host-side Quantity: the user-visible numpy arrays of one field.

Mirror of the parts of anuga/abstract_2d_finite_volumes/quantity.py the hot path
touches: the arrays (quantity.py:63-103), ``set_values`` for constants, arrays,
callables and other quantities at 'vertices' / 'centroids' (:709-1010,
:1103-1200), ``interpolate`` (:quantity.c:665-688) and
``extrapolate_first_order``.  File/geospatial/raster sources are out of scope
(SURVEY.md section 2, rows 7-9).

Arrays are allocated on first access so that a 16M-triangle domain does not pay
for (N,3) arrays nobody reads.
"""
import numpy as np

_LAZY = {
    "vertex_values": 3, "edge_values": 3, "explicit_update": 1, "semi_implicit_update": 1,
    "centroid_backup_values": 1,
}


class Quantity:
    def __init__(self, domain, name=None):
        self.domain = domain
        self.name = name
        N = domain.number_of_triangles
        self.N = N
        self.centroid_values = np.zeros(N, dtype=np.float64)
        self.boundary_values = np.zeros(domain.boundary_length, dtype=np.float64)
        self._arrays = {}
        self.beta = 1.0
        # host copy is newer than the device copy (set by the setters below)
        self.host_dirty = True

    def __len__(self):
        return self.N

    def __getstate__(self):
        state = dict(self.__dict__)
        state.pop("_fetch", None)          # bound method of the owning Domain: restored by Domain.__setstate__
        return state

    def __setstate__(self, state):
        self.__dict__.update(state)

    def __getattr__(self, name):
        if name in _LAZY:
            arrays = self.__dict__["_arrays"]
            if name not in arrays:
                N = self.__dict__["N"]
                k = _LAZY[name]
                arrays[name] = np.zeros((N, 3) if k == 3 else N, dtype=np.float64)
                if k == 3:   # first touch: consistent with the centroid values
                    arrays[name][:] = self.__dict__["centroid_values"][:, None]
            fetch = self.__dict__.get("_fetch")
            if fetch is not None:
                fetch(self, name, arrays[name])
            return arrays[name]
        raise AttributeError(name)

    # ------------------------------------------------------------------
    def set_values(self, numeric=None, quantity=None, function=None, location="vertices",
                   indices=None, expression=None, polygon=None, **unsupported):
        for k, v in unsupported.items():
            if v not in (None, False):
                raise NotImplementedError("set_values(%s=...) is outside the hot-path scope" % k)
        if expression is not None:                  # quantity.py:860-864
            assert numeric is None and quantity is None and function is None
            quantity = self.domain.create_quantity_from_expression(expression)
        if polygon is not None:
            # quantity.py:778-800: a constant over the triangles whose centroid lies in the polygon;
            # always applied at centroids, after which ALL vertex / edge values are first order
            if indices is not None:
                raise Exception("Only one of polygon and indices can be specified")
            if numeric is None or not isinstance(numeric, (float, int)):
                raise Exception("With polygon selected, set_quantity must provide the keyword numeric "
                                "and it must (currently) be a constant.")
            from .structures import Region
            indices = Region(self.domain, polygon=polygon).indices
            location = "centroids"
        if location == "edges":
            raise Exception("edges has been deprecated as valid location")
        if location not in ("vertices", "centroids", "unique vertices"):
            raise Exception("Invalid location: %s" % location)
        if location == "unique vertices":
            raise NotImplementedError("location='unique vertices' is outside the hot-path scope")
        given = [x for x in (numeric, quantity, function) if x is not None]
        if len(given) != 1:
            raise Exception("Exactly one of the arguments numeric, quantity, function must be present.")
        if function is not None:
            assert callable(function), "Argument function must be callable"
            numeric = function
        if quantity is not None:
            numeric = quantity

        if isinstance(numeric, Quantity):
            # set_values_from_quantity (quantity.py:1103-1117): always through the vertex values
            location = "vertices"
            self._from_array(np.array(numeric.vertex_values), location, indices)
        elif callable(numeric):
            self._from_function(numeric, location, indices)
        elif isinstance(numeric, (list, tuple, np.ndarray)):
            self._from_array(np.asarray(numeric, dtype=np.float64), location, indices)
        else:
            self._from_constant(float(numeric), location, indices)

        if location == "vertices":
            self.interpolate()
        else:
            self.extrapolate_first_order()
        self.host_dirty = True

    def _from_constant(self, X, location, indices):
        if location == "centroids":
            if indices is None:
                self.centroid_values[:] = X
            else:
                self.centroid_values[np.asarray(indices, dtype=np.int64)] = X
        else:
            if indices is None:
                self.vertex_values[:] = X
            else:
                self.vertex_values[np.asarray(indices, dtype=np.int64)] = X

    def _from_array(self, values, location, indices):
        N = self.N
        if location == "centroids":
            if indices is None:
                assert values.shape == (N,), "Number of values must match number of elements"
                self.centroid_values[:] = values
            else:
                self.centroid_values[np.asarray(indices, dtype=np.int64)] = values
        else:
            if values.ndim == 1:
                # one value per mesh node (quantity.py set_values_from_array, 1-D branch)
                assert indices is None
                assert values.shape[0] == self.domain.number_of_nodes, \
                    "1-D vertex arrays must hold one value per node"
                self.vertex_values[:] = values[self.domain.triangles]
            else:
                if indices is None:
                    assert values.shape == (N, 3), "Array must be N x 3"
                    self.vertex_values[:] = values
                else:
                    self.vertex_values[np.asarray(indices, dtype=np.int64)] = values

    def _from_function(self, f, location, indices):
        if location == "centroids":
            C = self.domain.centroid_coordinates
            if indices is not None:
                C = C[np.asarray(indices, dtype=np.int64)]
            res = f(C[:, 0], C[:, 1])
            if np.isscalar(res):
                self._from_constant(float(res), location, indices)
            else:
                self._from_array(np.asarray(res, dtype=np.float64), location, indices)
        else:
            V = self.domain.vertex_coordinates
            values = f(V[:, 0], V[:, 1])
            if np.isscalar(values):
                self._from_constant(float(values), location, indices)
                return
            values = np.asarray(values, dtype=np.float64).reshape(self.N, 3)
            if indices is None:
                self.vertex_values[:] = values
            else:
                idx = np.asarray(indices, dtype=np.int64)
                self.vertex_values[idx] = values[idx]

    # -- arithmetic (quantity.py:230-420): what expressions such as 'elevation + 0.05' evaluate with --
    def _as_quantity(self, other):
        if isinstance(other, Quantity):
            return other
        Q = Quantity(self.domain)
        Q.set_values(other)
        return Q

    def _triple(self, op, Q):
        result = Quantity(self.domain)
        result.vertex_values[:] = op(self.vertex_values, Q.vertex_values)
        result.edge_values[:] = op(self.edge_values, Q.edge_values)
        result.centroid_values[:] = op(self.centroid_values, Q.centroid_values)
        return result

    def __neg__(self):
        Q = Quantity(self.domain)
        Q.set_values(-self.vertex_values)
        return Q

    def __add__(self, other):
        Q = Quantity(self.domain)
        Q.set_values(other)
        result = Quantity(self.domain)
        result.set_values(self.vertex_values + Q.vertex_values)
        return result

    __radd__ = __add__

    def __sub__(self, other):
        return self + -other

    def __rsub__(self, other):
        return -self + other

    def __mul__(self, other):
        return self._triple(lambda a, b: a * b, self._as_quantity(other))

    __rmul__ = __mul__

    def __truediv__(self, other):
        eps = 1.0e-12                               # anuga/config.py:12 (safe division)
        return self._triple(lambda a, b: a / (b + eps), self._as_quantity(other))

    def __pow__(self, other):
        assert not isinstance(other, Quantity)
        result = Quantity(self.domain)
        result.vertex_values[:] = self.vertex_values ** other
        result.edge_values[:] = self.edge_values ** other
        result.centroid_values[:] = self.centroid_values ** other
        return result

    def maximum(self, other):
        Q = self._as_quantity(other)
        self.vertex_values[:] = np.maximum(self.vertex_values, Q.vertex_values)
        self.edge_values[:] = np.maximum(self.edge_values, Q.edge_values)
        self.centroid_values[:] = np.maximum(self.centroid_values, Q.centroid_values)
        self.host_dirty = True
        return self

    def minimum(self, other):
        Q = self._as_quantity(other)
        self.vertex_values[:] = np.minimum(self.vertex_values, Q.vertex_values)
        self.edge_values[:] = np.minimum(self.edge_values, Q.edge_values)
        self.centroid_values[:] = np.minimum(self.centroid_values, Q.centroid_values)
        self.host_dirty = True
        return self

    # ------------------------------------------------------------------
    def interpolate(self):
        """centroid = mean of vertices, edges = vertex mid-points (quantity.c:665-688)"""
        v = self.vertex_values
        q0, q1, q2 = v[:, 0], v[:, 1], v[:, 2]
        self.centroid_values[:] = (q0 + q1 + q2) / 3.0
        e = self.edge_values
        e[:, 0] = 0.5 * (q1 + q2)
        e[:, 1] = 0.5 * (q0 + q2)
        e[:, 2] = 0.5 * (q0 + q1)

    def extrapolate_first_order(self):
        c = self.centroid_values
        if "vertex_values" in self._arrays:
            self._arrays["vertex_values"][:] = c[:, None]
        if "edge_values" in self._arrays:
            self._arrays["edge_values"][:] = c[:, None]

    def get_values(self, location="vertices", indices=None):
        if location == "centroids":
            a = self.centroid_values
        elif location == "edges":
            a = self.edge_values
        elif location == "vertices":
            a = self.vertex_values
        else:
            raise NotImplementedError(location)
        return a if indices is None else a[np.asarray(indices, dtype=np.int64)]

    def get_vertex_values(self, xy=True, smooth=None, centroid_averaging=None, precision=None):
        """One value per mesh node (smooth: mean over the triangles sharing the node of their vertex
        values or - centroid_averaging, the DE algorithms' setting - of their centroid values, summed
        in the reference's order) or per triangle vertex, optionally with coordinates and
        connectivity (quantity.py:2158-2228, quantity.c:823-910)."""
        d = self.domain
        if smooth is None:
            smooth = bool(getattr(d, "smooth", False))
        if centroid_averaging is None:
            centroid_averaging = bool(getattr(d, "using_centroid_averaging", False))
        if precision is None:
            precision = np.float64
        if smooth:
            count, order = d.mesh.build_inverted_triangle_structure()
            V = d.triangles
            if centroid_averaging:
                vals = self.centroid_values[order // 3]
            else:
                vals = self.vertex_values.reshape(-1)[order]
            start = np.concatenate([[0], np.cumsum(count)[:-1]])
            A = np.zeros(d.number_of_nodes, dtype=np.float64)
            for j in range(int(count.max())):           # sequential sums, like the C loop
                m = count > j
                A[m] += vals[start[m] + j]
            A = (A / count).astype(precision)
            points = d.nodes
        else:
            V = np.arange(3 * self.N, dtype=np.int64).reshape(-1, 3)
            points = d.vertex_coordinates
            A = self.vertex_values.flatten().astype(precision)
        if xy:
            return points[:, 0].astype(precision), points[:, 1].astype(precision), A, V
        return A, V

    # -- extrema on centroids (quantity.py:1796-1930): ties go to the first cell ----------------------------
    def get_extremum_index(self, mode=None, indices=None):
        V = self.get_values(location="centroids", indices=indices)
        if mode is None or mode == "max":
            i = np.argmax(V)
        elif mode == "min":
            i = np.argmin(V)
        else:
            raise ValueError("Bad mode value, got: %s" % str(mode))
        return i if indices is None else indices[i]

    def get_maximum_index(self, indices=None):
        return self.get_extremum_index(mode="max", indices=indices)

    def get_maximum_value(self, indices=None):
        return self.get_values(location="centroids")[self.get_maximum_index(indices)]

    def get_maximum_location(self, indices=None):
        x, y = self.domain.get_centroid_coordinates()[self.get_maximum_index(indices)]
        return x, y

    def get_minimum_index(self, indices=None):
        return self.get_extremum_index(mode="min", indices=indices)

    def get_minimum_value(self, indices=None):
        return self.get_values(location="centroids")[self.get_minimum_index(indices)]

    def get_minimum_location(self, indices=None):
        x, y = self.domain.get_centroid_coordinates()[self.get_minimum_index(indices)]
        return x, y

    def get_integral(self, full_only=True):
        areas = self.domain.areas
        if full_only:
            m = self.domain.tri_full_flag == 1
            return float(np.sum(areas[m] * self.centroid_values[m]))
        return float(np.sum(areas * self.centroid_values))
