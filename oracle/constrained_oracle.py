"""Oracle: size-constrained (semi-supervised) k-means of SCD, restated on torch-CPU / NumPy.

TEST INFRASTRUCTURE - not shipped, never imported by ``scd_b200``.

Follows ``local_utils/sskm_constrained.py`` (paths relative to the reference checkout):
  * ``K_Means`` :15-187 (``fit_once`` :47-85, ``fit_mix_once`` :87-139, ``fit`` :141-163, ``fit_mix`` :165-187)
  * ``_labels_constrained`` :226-274, ``minimum_cost_flow_problem_graph`` :277-328,
    ``solve_min_cost_flow_graph`` :331-356.

**Parity unpinned for the solver.**  The flow problem is solved in the reference by Google OR-Tools 9.3.10497
(``requirements.txt:103``; ``ortools.graph.pywrapgraph.SimpleMinCostFlow``, a cost-scaling push-relabel in
C++), which is neither in the reference tree nor installed here.  ``StandInMinCostFlow`` below solves the very
same explicit graph (same arcs, capacities, unit costs, supplies) as a linear programme with SciPy's HiGHS -
the constraint matrix of a flow problem is totally unimodular, so a vertex optimum is integral - and returns
the per-arc flows through the ``SimpleMinCostFlowVectorized`` interface the reference calls.  What *is* pinned
against the real reference code (``oracle/gen_golden.py`` imports ``local_utils/sskm_constrained.py`` with only
the OR-Tools module replaced by this stand-in): the graph arrays of ``minimum_cost_flow_problem_graph`` and the
whole ``fit`` / ``fit_mix`` plumbing around the solver.  Optimal flows are not unique under tied integer costs,
so the CUDA path is compared by optimal total cost + capacity feasibility, and by labels only where the
optimum is unique (continuous random data).
"""
from __future__ import annotations

import numpy as np
import torch
from sklearn.utils import check_random_state

from . import kmeans_oracle


class StandInMinCostFlow:
    """The five members of ``SimpleMinCostFlowVectorized`` the reference uses (:333-353), over SciPy HiGHS."""

    OPTIMAL = 1
    INFEASIBLE = 3

    def __init__(self):
        self._tail = self._head = self._cap = self._cost = None
        self._supply = None
        self._flow = None

    def AddArcWithCapacityAndUnitCostVectorized(self, tail, head, capacity, unit_cost):
        self._tail, self._head = np.asarray(tail, dtype=np.int64), np.asarray(head, dtype=np.int64)
        self._cap, self._cost = np.asarray(capacity, dtype=np.float64), np.asarray(unit_cost, dtype=np.float64)

    def SetNodeSupplyVectorized(self, node, supply):
        s = np.zeros(int(np.max(node)) + 1, dtype=np.float64)
        s[np.asarray(node, dtype=np.int64)] = np.asarray(supply, dtype=np.float64)
        self._supply = s

    def Solve(self):
        from scipy.optimize import linprog
        from scipy.sparse import coo_matrix
        n_arcs, n_nodes = len(self._tail), len(self._supply)
        arcs = np.arange(n_arcs)
        # conservation: outflow - inflow = supply at every node
        A = coo_matrix((np.concatenate([np.ones(n_arcs), -np.ones(n_arcs)]),
                        (np.concatenate([self._tail, self._head]), np.concatenate([arcs, arcs]))),
                       shape=(n_nodes, n_arcs)).tocsr()
        res = linprog(self._cost, A_eq=A, b_eq=self._supply, bounds=np.stack([np.zeros(n_arcs), self._cap], axis=1),
                      method='highs-ds')
        if res.status != 0:
            return self.INFEASIBLE
        self._flow = np.rint(res.x).astype(np.int32)
        self.total_cost = int(np.rint(res.fun))
        return self.OPTIMAL

    def FlowVectorized(self, arc):
        return self._flow[np.asarray(arc, dtype=np.int64)]


def minimum_cost_flow_problem_graph(X, C, D, size_min, size_max):
    """ref :277-328 - node ids X [0, n_X), C' [n_X, n_X+n_C), C [n_X+n_C, n_X+2 n_C), artificial n_X+2 n_C."""
    n_X, n_C = X.shape[0], C.shape[0]
    X_ix = np.arange(n_X)
    C_dummy_ix = np.arange(n_X, n_X + n_C)                                          # :287
    C_ix = np.arange(n_X + n_C, n_X + 2 * n_C)                                      # :288
    art_ix = n_X + 2 * n_C                                                          # :289
    edges_X_C_dummy = np.stack([np.repeat(X_ix, n_C), np.tile(C_dummy_ix, n_X)], axis=1)   # :292 cartesian, X-major
    edges_C_dummy_C = np.stack([C_dummy_ix, C_ix], axis=1)                          # :293
    edges_C_art = np.stack([C_ix, art_ix * np.ones(n_C)], axis=1)                   # :294
    edges = np.concatenate([edges_X_C_dummy, edges_C_dummy_C, edges_C_art])         # :296
    costs_X_C_dummy = D.reshape(D.size)                                             # :299
    costs = np.concatenate([costs_X_C_dummy, np.zeros(edges.shape[0] - len(costs_X_C_dummy))])   # :300
    capacities = np.concatenate([np.ones(edges_X_C_dummy.shape[0]), size_max * np.ones(n_C), n_X * np.ones(n_C)])  # :303-309
    supplies = np.concatenate([np.ones(n_X), np.zeros(n_C), -1 * size_min * np.ones(n_C),
                               [-1 * (n_X - n_C * size_min)]])                      # :312-320
    edges = edges.astype('int32')                                                   # :323
    costs = np.around(costs * 1000, 0).astype('int32')                              # :324
    capacities = capacities.astype('int32')
    supplies = supplies.astype('int32')
    return edges, costs, capacities, supplies, n_C, n_X


def solve_min_cost_flow_graph(edges, costs, capacities, supplies, n_C, n_X, solver_cls=StandInMinCostFlow):
    """ref :331-356 with the solver class injected."""
    mcf = solver_cls()
    if (edges.dtype != 'int32') or (costs.dtype != 'int32') or (capacities.dtype != 'int32') or (supplies.dtype != 'int32'):
        raise ValueError("`edges`, `costs`, `capacities`, `supplies` must all be int dtype")          # :335-337
    mcf.AddArcWithCapacityAndUnitCostVectorized(edges[:, 0], edges[:, 1], capacities, costs)        # :343
    mcf.SetNodeSupplyVectorized(np.arange(len(supplies), dtype='int32'), supplies)                  # :346
    if mcf.Solve() != mcf.OPTIMAL:                                                                  # :349-350
        raise Exception('There was an issue with the min cost flow input.')
    labels_M = mcf.FlowVectorized(np.arange(n_X * n_C, dtype='int32')).reshape(n_X, n_C)            # :353
    return labels_M.argmax(axis=1)                                                                  # :355


def labels_constrained(X, centers, D_sqrt, size_min, size_max):
    """ref :226-274 -> ``(labels int32 [N], inertia float32)``; ``D_sqrt`` = ``torch.sqrt(dist).numpy()`` (:116)."""
    edges, costs, capacities, supplies, n_C, n_X = minimum_cost_flow_problem_graph(X, centers, D_sqrt, size_min, size_max)
    labels = solve_min_cost_flow_graph(edges, costs, capacities, supplies, n_C, n_X).astype(np.int32)   # :262-266
    distances = D_sqrt[np.arange(D_sqrt.shape[0]), labels] ** 2                                        # :271
    return labels, distances.sum()                                                                     # :272


def int_costs(D_sqrt: np.ndarray) -> np.ndarray:
    """The X -> C' arc costs (:299 + :324)."""
    return np.around(D_sqrt * 1000, 0).astype('int32')


def optimal_total_cost(cost: np.ndarray, size_min: int, size_max: int):
    """Optimal objective of the assignment problem on an int cost matrix (None if infeasible) through the same
    explicit graph + stand-in solver."""
    n, k = cost.shape
    dummy = np.zeros((n, 1)), np.zeros((k, 1))
    edges, _costs, capacities, supplies, n_C, n_X = minimum_cost_flow_problem_graph(dummy[0], dummy[1], np.zeros((n, k)), size_min, size_max)
    costs = np.concatenate([cost.reshape(-1), np.zeros(2 * k)]).astype('int32')
    mcf = StandInMinCostFlow()
    mcf.AddArcWithCapacityAndUnitCostVectorized(edges[:, 0], edges[:, 1], capacities, costs)
    mcf.SetNodeSupplyVectorized(np.arange(len(supplies), dtype='int32'), supplies)
    if mcf.Solve() != mcf.OPTIMAL:
        return None
    return mcf.total_cost


class K_Means(kmeans_oracle.K_Means):
    """ref :15-187; ``kpp`` (:28-44) is the local copy's (IndexError when no candidate)."""

    def __init__(self, k=3, tolerance=1e-4, max_iterations=100, size_min=100, size_max=1000, init='k-means++', n_init=10,
                 random_state=None, n_jobs=None, pairwise_batch_size=None):
        super().__init__(k=k, tolerance=tolerance, max_iterations=max_iterations, init=init, n_init=n_init,
                         random_state=random_state, n_jobs=n_jobs, pairwise_batch_size=pairwise_batch_size)
        self.size_min, self.size_max = size_min, size_max

    def _assign(self, X, centers):
        dist = kmeans_oracle.pairwise_distance(X, centers, self.pairwise_batch_size)                # :66 / :115
        return labels_constrained(X.cpu().numpy(), centers.cpu().numpy(), torch.sqrt(dist).cpu().numpy(),
                                  self.size_min, self.size_max)                                      # :67 / :116

    def fit_once(self, X, random_state):                                                             # :47-85
        centers = torch.zeros(self.k, X.shape[1]).type_as(X)
        if self.init == 'k-means++':
            centers = self.kpp(X, k=self.k, random_state=random_state)
        elif self.init == 'random':
            rs = check_random_state(self.random_state)
            idx = rs.choice(len(X), self.k, replace=False)
            for i in range(self.k):
                centers[i] = X[idx[i]]
        else:
            for i in range(self.k):
                centers[i] = X[i]
        best_labels = best_inertia = best_centers = None
        n_done = 0
        for it in range(self.max_iterations):
            n_done = it + 1
            centers_old = centers.clone()
            labels, inertia = self._assign(X, centers)
            labels = torch.from_numpy(labels)                                                        # :68 (int32, CPU)
            kmeans_oracle.mstep(X, labels, centers)                                                  # :71-74
            if best_inertia is None or inertia < best_inertia:                                       # :76-79
                best_labels, best_centers, best_inertia = labels.clone(), centers.clone(), inertia
            if kmeans_oracle.center_shift(centers, centers_old) ** 2 < self.tolerance:               # :81-84
                break
        return best_labels, best_inertia, best_centers, n_done

    def fit_mix_once(self, u_feats, l_feats, l_targets, random_state):                               # :87-139
        l_classes = torch.unique(l_targets)
        l_centers = torch.stack([l_feats[l_targets.eq(c).nonzero().squeeze(1)].mean(0) for c in l_classes])
        cat_feats = torch.cat((l_feats, u_feats))
        labels = -torch.ones(len(cat_feats)).type_as(cat_feats).long()
        classes_np = l_classes.cpu().long().numpy()
        targets_np = l_targets.cpu().long().numpy()
        l_num = len(targets_np)
        remap = {cid: new for new, cid in enumerate(classes_np)}
        i = None
        for i in range(l_num):
            labels[i] = remap[targets_np[i]]
        centers = self.kpp(u_feats, l_centers, k=self.k, random_state=random_state)                  # :108
        best_labels = best_inertia = best_centers = None
        for _it in range(self.max_iterations):
            centers_old = centers.clone()
            u_labels, u_inertia = self._assign(u_feats, centers)                                     # :115-117
            u_labels = torch.from_numpy(u_labels).type_as(labels)
            l_mindist = torch.sum((l_feats - centers[labels[:l_num]]) ** 2, dim=1)                   # :120
            inertia = u_inertia + l_mindist.sum()                                                    # :121-122
            labels[l_num:] = u_labels                                                                # :123
            kmeans_oracle.mstep(cat_feats, labels, centers)                                          # :125-128
            if best_inertia is None or inertia < best_inertia:                                       # :130-133
                best_labels, best_centers, best_inertia = labels.clone(), centers.clone(), inertia
            if kmeans_oracle.center_shift(centers, centers_old) ** 2 < self.tolerance:               # :135-138
                break
        return best_labels, best_inertia, best_centers, i + 1                                        # stale ``i`` as in :139
