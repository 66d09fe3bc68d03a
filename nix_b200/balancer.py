"""Rank boundaries from per-chunk loads: the assignment of the reference's Balancer, for hosts that are not a
nix application (bench.py, the tests).  A nix application keeps using its own nix::Balancer unchanged; this
module restates balancer.cpp so that the Python host hands nixb200_domain_rebalance the SAME boundaries:

    assign_binarysearch   balancer.cpp:74-99     boundary[i] = last index with cumload <= i * mean
    assign_smilei         balancer.cpp:8-72      one relaxation sweep (Derouillat et al. 2018), every boundary
                                                 moves towards its target but never past its old neighbours
    assign_initial        balancer.cpp:101-124   binary search, or (if that leaves a rank empty) smilei sweeps
                                                 from the uniform assignment
    assign                balancer.cpp:126-132   one smilei sweep from the current boundaries

Pinned against the reference's own class through tests/golden/balancer.npz (tests/golden/make_balancer_golden.py).
"""
import bisect


def _cumload(load):
    cum = [0.0]
    for v in load:
        cum.append(cum[-1] + float(v))
    return cum


def assign_smilei(load, boundary):
    load = [float(v) for v in load]
    boundary = [int(b) for b in boundary]
    nr = len(boundary) - 1
    cum = _cumload(load)
    mean = cum[-1] / nr
    old = list(boundary)
    for i in range(1, nr):
        target = mean * i
        current = cum[boundary[i]]
        if current > target:  # possibly move the boundary backward
            index = boundary[i] - 1
            while abs(current - target) > abs(current - target - load[index]):
                current -= load[index]
                index -= 1
            boundary[i] = index + 1 if index >= old[i - 1] else old[i - 1] + 1
        else:  # move it forward
            index = boundary[i]
            while abs(current - target) > abs(current - target + load[index]):
                current += load[index]
                index += 1
            boundary[i] = index if index < old[i + 1] else old[i + 1] - 1
    return boundary, boundary != old


def assign_binarysearch(load, nrank):
    cum = _cumload(load)
    mean = cum[-1] / nrank
    boundary = [0] * (nrank + 1)
    boundary[nrank] = len(load)
    for i in range(1, nrank):
        boundary[i] = bisect.bisect_right(cum, mean * i) - 1
    ok = boundary[0] == 0 and all(boundary[i + 1] > boundary[i] for i in range(1, nrank))
    return boundary, ok


def assign_initial(load, nrank):
    boundary, ok = assign_binarysearch(load, nrank)
    if not ok:
        boundary, _ = assign_binarysearch([1.0] * len(load), nrank)
        for _ in range(100):
            boundary, changed = assign_smilei(load, boundary)
            if not changed:
                break
    return boundary


def assign(load, boundary):
    return assign_smilei(load, boundary)[0]
