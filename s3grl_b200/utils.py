"""Drop-in mirror of the reference's dispatcher `extract_enclosing_subgraphs`
(utils.py:446-496): same signature, same branch order on `sign_kwargs` / `powers_of_A`.  The keyword-only
`device`, `output_device` and `graph` are passed through to OptimizedSignOperations (tuned_sign.py)."""
import torch

from .data import PrecomputedList
from .tuned_sign import OptimizedSignOperations


def extract_enclosing_subgraphs(link_index, A, x, y, num_hops, node_label='drnl',
                                ratio_per_hop=1.0, max_nodes_per_hop=None,
                                directed=False, A_csc=None, rw_kwargs=None, sign_kwargs=None, powers_of_A=None,
                                data=None, *, device=None, output_device=None, graph=None, cap_seed=None):
    where = dict(device=device, output_device=output_device, graph=graph)
    pos_where = dict(where, cap_seed=cap_seed)      # the PoS flows take the per-hop caps (utils.py:66-70); SoP has no BFS
    if not sign_kwargs:
        raise NotImplementedError("only the SIGN flows (sign_kwargs) are on the accelerated path; "
                                  "SEAL / DRNL subgraph datasets are out of scope (SURVEY.md §2)")
    if powers_of_A and sign_kwargs['optimize_sign'] and sign_kwargs['sign_type'] == 'hybrid':
        # reference utils.py:454-480: PoS x, x1..xK then SoP x2..xK appended as x{K+1}..x{2K-1}
        sign_k = sign_kwargs['sign_k']
        sup = OptimizedSignOperations.get_PoS_prepped_ds(link_index, num_hops, A, ratio_per_hop,
                                                         max_nodes_per_hop, directed, A_csc, x, y,
                                                         sign_kwargs, rw_kwargs, **pos_where)
        if sign_k == 1:
            return sup
        sop = OptimizedSignOperations.get_SoP_prepped_ds(powers_of_A, link_index, A, x, y, **where)
        return PrecomputedList(sup.xs + [t.to(sup.xs[0].device) for t in sop.xs[2:sign_k + 1]], sup.row_ptr, sup.y)
    elif powers_of_A and sign_kwargs['optimize_sign']:
        return OptimizedSignOperations.get_SoP_prepped_ds(powers_of_A, link_index, A, x, y, **where)
    elif not powers_of_A and sign_kwargs['optimize_sign'] and not sign_kwargs['k_heuristic']:
        return OptimizedSignOperations.get_PoS_prepped_ds(link_index, num_hops, A, ratio_per_hop,
                                                          max_nodes_per_hop, directed, A_csc, x, y,
                                                          sign_kwargs, rw_kwargs, **pos_where)
    elif not powers_of_A and sign_kwargs['optimize_sign'] and sign_kwargs['k_heuristic']:
        return OptimizedSignOperations.get_PoS_Plus_prepped_ds(link_index, num_hops, A, ratio_per_hop,
                                                               max_nodes_per_hop, directed, A_csc, x, y,
                                                               sign_kwargs, rw_kwargs, **pos_where)
    elif not sign_kwargs['optimize_sign']:
        # SIGN + SEAL flow (reference utils.py:497-550)
        if powers_of_A:
            # reference utils.py:521-548 runs k_hop_subgraph on every global power with its own node set
            # and stacks operators whose row counts differ per power: there is no consistent layout to
            # reproduce, and no config uses it
            raise NotImplementedError("the non-optimised SoP flow (optimize_sign=False with powers_of_A, "
                                      "utils.py:521-548) is not supported")
        return OptimizedSignOperations.get_PoS_full_ds(link_index, num_hops, A, ratio_per_hop, max_nodes_per_hop,
                                                       directed, A_csc, x, y, sign_kwargs, rw_kwargs, node_label, **pos_where)
    else:
        raise NotImplementedError("No matching configuration for model data prep found. Please check code.")
