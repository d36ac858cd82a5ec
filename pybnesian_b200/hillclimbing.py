"""Greedy hill climbing (learning/algorithms/hillclimbing.{hpp,cpp}; pybindings_algorithms.cpp:75-234).

`estimate_hc` is restated statement for statement (same clones, same stop rules, same tabu / patience
logic, callbacks at iteration 0, after every accepted operator and at the end); the scores behind
`op_set.cache_scores` / `update_scores` are evaluated in batches on the GPU (operators.py, scores.py).
"""
import sys
import time

from .dataset import DataFrame
from .factors import FactorType
from .operators import ArcOperatorSet, ChangeNodeTypeSet, LocalScoreCache, OperatorPool, OperatorSet, OperatorTabuSet
from .scores import BIC, CVLikelihood, HoldoutLikelihood, Score, ValidatedLikelihood, ValidatedScore

MACHINE_TOL = 1.4901161193847656e-08  # util::machine_tol = sqrt(DBL_EPSILON) (util/math_constants.hpp:30)
INT_MAX = 2147483647


class Callback:
    """learning/algorithms/callbacks/callback.hpp: call(model, operator, score, iteration)."""

    def call(self, model, operator, score, iteration):
        raise NotImplementedError


class SaveModel(Callback):
    """learning/algorithms/callbacks/save_model.hpp:8-22: saves the model of every iteration as
    `folder_name/NNNNNN.pickle` (six digits, without the CPDs)."""

    def __init__(self, folder_name):
        self._folder_name = folder_name

    def call(self, model, operator, score, iteration):
        model.save("%s/%06d" % (self._folder_name, iteration), False)


def _validation_delta_score(model, val_score, variables, current_local_scores):
    """hillclimbing.hpp:46-60."""
    prev = 0.0
    nnew = 0.0
    for n in variables:
        prev += current_local_scores.local_score(model, n)
        current_local_scores.update_vlocal_score(model, val_score, n)
        nnew += current_local_scores.local_score(model, n)
    return nnew - prev


def _validate_type_restrictions(model, type_blacklist, type_whitelist):
    for lst, name in ((type_whitelist, "whitelist"), (type_blacklist, "blacklist")):
        for n, _ in lst:
            if not model.contains_node(n):
                raise ValueError("Node in the " + name + " (" + n + "), not present in the model.")
    white = {}
    for n, t in type_whitelist:
        white[n] = t
    for n, t in type_blacklist:
        if n in white and white[n] == t:
            raise ValueError("Node type " + str(t) + " for node " + n + " in blacklist and whitelist")


class GreedyHillClimbing:
    """pybnesian.GreedyHillClimbing."""

    last_run = None  # timing / counters of the last estimate() (not in the reference: bench instrumentation)

    def estimate(self, operators, score, start, arc_blacklist=(), arc_whitelist=(), type_blacklist=(), type_whitelist=(),
                 callback=None, max_indegree=0, max_iters=INT_MAX, epsilon=0, patience=0, verbose=0):
        arc_blacklist, arc_whitelist = list(arc_blacklist), list(arc_whitelist)
        type_blacklist, type_whitelist = list(type_blacklist), list(type_whitelist)
        # estimate_checks (hillclimbing.hpp:274-304)
        if not score.compatible_bn(start):
            raise ValueError("BayesianNetwork is not compatible with the score.")
        from .operators import _validate_restrictions
        _validate_restrictions(start, arc_blacklist, arc_whitelist)
        _validate_type_restrictions(start, type_blacklist, type_whitelist)
        validated = isinstance(score, ValidatedScore)
        zero_patience = patience == 0
        op_set = operators

        current_model = start.clone()
        current_model.force_type_whitelist(type_whitelist)
        if current_model.has_unknown_node_types():
            score_data = score.data()
            if DataFrame.wrap(score_data).num_columns == 0:
                raise ValueError("The score does not have data to detect the node types. Set the node types for all the "
                                 "nodes in the Bayesian network or use an score that uses data (it implements Score::data).")
            if not DataFrame.wrap(score_data).has_columns(current_model.nodes()):
                raise ValueError("The score data does not contain all the nodes of the model.")
            current_model.set_unknown_node_types(score_data, type_blacklist)
        current_model.check_blacklist(arc_blacklist)
        current_model.force_whitelist(arc_whitelist)

        op_set.set_arc_blacklist(arc_blacklist)
        op_set.set_arc_whitelist(arc_whitelist)
        op_set.set_type_blacklist(type_blacklist)
        op_set.set_type_whitelist(type_whitelist)
        op_set.set_max_indegree(max_indegree)

        prev_current_model = current_model.clone()
        best_model = current_model

        local_validation = None
        if validated:
            local_validation = LocalScoreCache(current_model)
            local_validation.cache_vlocal_scores(current_model, score)

        t0 = time.perf_counter()
        op_set.cache_scores(current_model, score)
        t_cache = time.perf_counter() - t0
        p = 0
        accumulated_offset = 0.0
        tabu_set = OperatorTabuSet()
        applied = []
        iter_times = []

        if callback is not None:
            callback.call(current_model, None, score, 0)

        it = 0
        while it < max_iters:
            it += 1
            t_it = time.perf_counter()
            best_op = op_set.find_max(current_model) if zero_patience else op_set.find_max(current_model, tabu_set)
            if best_op is None or (best_op.delta() - epsilon) < MACHINE_TOL:
                break
            best_op.apply(current_model)
            nodes_changed = best_op.nodes_changed(current_model)
            if validated:
                validation_delta = _validation_delta_score(current_model, score, nodes_changed, local_validation)
            else:
                validation_delta = best_op.delta()

            if (validation_delta + accumulated_offset) > MACHINE_TOL:
                if not zero_patience:
                    if p > 0:
                        best_model = current_model
                        p = 0
                        accumulated_offset = 0.0
                    tabu_set.clear()
            else:
                if zero_patience:
                    best_model = prev_current_model
                    break
                else:
                    if p == 0:
                        best_model = prev_current_model.clone()
                    p += 1
                    if p > patience:
                        break
                    accumulated_offset += validation_delta
                    tabu_set.insert(best_op.opposite(current_model))

            best_op.apply(prev_current_model)
            applied.append(best_op)
            if callback is not None:
                callback.call(current_model, best_op, score, it)
            op_set.update_scores(current_model, score, nodes_changed)
            iter_times.append(time.perf_counter() - t_it)
            if verbose:
                msg = str(best_op) + (" | Validation delta: %f" % validation_delta if validated else "")
                print(msg, file=sys.stderr)

        op_set.finished()
        if callback is not None:
            callback.call(best_model, None, score, it)
        self.last_run = {"operators": applied, "cache_scores_s": t_cache, "iteration_s": iter_times, "iterations": it}
        return best_model


def _check_valid_score(df, bn_type, score, seed, num_folds, test_holdout_ratio):
    """util::check_valid_score (util/validate_options.cpp:16-49); bic / bge are not on this path."""
    from .models import GaussianNetworkType, KDENetworkType, SemiparametricBNType
    if score is not None:
        if score == "cv-lik":
            return CVLikelihood(df, num_folds, seed)
        if score == "holdout-lik":
            return HoldoutLikelihood(df, test_holdout_ratio, seed)
        if score == "validated-lik":
            return ValidatedLikelihood(df, test_holdout_ratio, num_folds, seed)
        if score == "bic":
            return BIC(df)
        if score == "bge":
            raise NotImplementedError("score \"bge\" is outside the likelihood-score path of pybnesian_b200")
        raise ValueError("Wrong Bayesian Network score \"" + score + "\" specified. The possible alternatives are "
                         "\"bic\" (Bayesian Information Criterion), \"bge\" (Bayesian Gaussian equivalent), "
                         "\"cv-lik\" (Cross-Validated likelihood), \"holdout-l\" (Hold-out likelihood) "
                         " or \"validated-lik\" (Validated likelihood with cross-validation).")
    if bn_type == SemiparametricBNType() or bn_type == KDENetworkType():
        return ValidatedLikelihood(df, test_holdout_ratio, num_folds, seed)
    if bn_type == GaussianNetworkType():
        return BIC(df)
    raise ValueError("Default score not defined for " + str(bn_type) + ".")


def _check_valid_operators(bn_type, operators, arc_blacklist, arc_whitelist, max_indegree, type_whitelist):
    """util::check_valid_operators (util/validate_options.cpp:51-91).  As in the reference, the type
    whitelist lands in ChangeNodeTypeSet's first (blacklist) parameter."""
    from .models import GaussianNetworkType, KDENetworkType, SemiparametricBNType
    res = []
    if operators:
        for op in operators:
            if op == "arcs":
                res.append(ArcOperatorSet(arc_blacklist, arc_whitelist, max_indegree))
            elif op == "node_type":
                if bn_type != SemiparametricBNType():
                    raise ValueError("Operator \"node_type\" is not compabible with Bayesian network type \"" + str(bn_type) + "\"")
                res.append(ChangeNodeTypeSet(type_whitelist))
            else:
                raise ValueError("Wrong operator set \"" + op + "\". Valid choices are:\"arcs\" (Changes in arcs; addition, "
                                 "removal and flip) or \"node_type\" (Change of node type)")
    else:
        if bn_type == GaussianNetworkType() or bn_type == KDENetworkType():
            res.append(ArcOperatorSet(arc_blacklist, arc_whitelist, max_indegree))
        elif bn_type == SemiparametricBNType():
            res.append(ArcOperatorSet(arc_blacklist, arc_whitelist, max_indegree))
            res.append(ChangeNodeTypeSet(type_whitelist))
        else:
            raise ValueError("Default operators not defined for " + str(bn_type) + ".")
    return res[0] if len(res) == 1 else OperatorPool(res)


def hc(df, bn_type=None, start=None, score=None, operators=None, arc_blacklist=(), arc_whitelist=(), type_blacklist=(),
       type_whitelist=(), callback=None, max_indegree=0, max_iters=INT_MAX, epsilon=0, patience=0, seed=None,
       num_folds=10, test_holdout_ratio=0.2, verbose=0):
    """pybnesian.hc (learning/algorithms/hillclimbing.cpp:26-90)."""
    from .dataset import _random_seed
    if bn_type is None and start is None:
        raise ValueError("\"bn_type\" or \"start\" parameter must be specified.")
    iseed = _random_seed(seed)
    bn_type_ = start.type() if start is not None else bn_type
    ops = _check_valid_operators(bn_type_, operators, list(arc_blacklist), list(arc_whitelist), max_indegree, list(type_whitelist))
    if max_iters == 0:
        max_iters = INT_MAX
    frame = DataFrame.wrap(df)
    start_model = start if start is not None else bn_type_.new_bn(frame.columns)
    sc = _check_valid_score(frame, bn_type_, score, iseed, num_folds, test_holdout_ratio)
    return GreedyHillClimbing().estimate(ops, sc, start_model, arc_blacklist, arc_whitelist, type_blacklist, type_whitelist,
                                         callback, max_indegree, max_iters, epsilon, patience, verbose)
