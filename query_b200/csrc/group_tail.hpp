// group_tail.hpp — the operators the reference runs on FinalGroup's output, restated over the (tiny) group result:
//   Let (LETTING) -> Filter (HAVING) -> InitialProject -> [FinalProject] -> Order -> Offset -> Limit -> [FinalProject]
// (SURVEY.md 8f rows 1-2).  A million documents have become a handful of groups by now, so this is host code: what it
// saves is the boxing of every group into Go values before rows are dropped by HAVING / LIMIT.
//
//   planner/build_select_sub.go:217-235,276-296 (operator order)      planner/build_select.go:75-110 (Order/Offset/Limit)
//   execution/let.go:50-62          execution/filter.go:49-61          execution/project_initial.go:52-144
//   execution/project_final.go:51-59   execution/order.go:50-170      execution/offset.go:53-83   execution/limit.go:53-85
//   algebra/result.go:358-374 (implicit aliases)   algebra/aggregate.go:97-118 (aggregate lookup by text)
//   value/object.go:246-255 (a MISSING projection value leaves the field out)
#pragma once
#include <memory>
#include <string>
#include <vector>

#include "expr.hpp"
#include "json.hpp"
#include "query.hpp"

namespace n1 {
namespace execution {

// Value.Collate over scalar values (value/value.go:69-79 type order, integer.go:100-130, float.go:123-172, string.go:116-126)
int collate_values(const HValue& a, const HValue& b);
// Evaluates a constant expression (OFFSET / LIMIT operands); INELIGIBLE when it references anything.
HValue eval_constant(const Expr& e);
std::string value_to_json(const HValue& v);

struct TailTerm { std::string expr, as, alias; ExprP e; };
struct TailSort { std::string expr; bool desc = false; ExprP e; };
struct TailBinding { std::string var, expr; ExprP e; };

class GroupTail {
  public:
    std::vector<TailBinding> letting;
    std::string having;
    ExprP having_e;
    bool has_project = false;
    std::vector<TailTerm> terms;
    std::vector<TailSort> order;
    bool has_offset = false, has_limit = false;
    i64 offset = 0, limit = 0;
    bool final_project = false;
    int inner_consumed = 0;  // operators taken from the chain's own Sequence (after FinalGroup)
    int outer_consumed = 0;  // operators taken from the enclosing Sequence (Order, Offset, Limit, FinalProject)
    std::vector<std::string> operators;  // names, in execution order (for MarshalJSON / tests)
    std::string keyspace_alias;          // the item's own field: an explicit projection alias must not shadow it
    // SELECT DISTINCT: the operator groups by the projected terms, so its groups ARE the distinct
    // rows (execution/distinct.go:60-72 keys its set by the projection object); the InitialProject may then carry
    // "distinct": true and the Distinct operators that follow it are absorbed
    bool distinct_by_keys = false;

    bool empty() const { return operators.empty(); }
    // Adds one plan operator (already parsed JSON); returns false - and adds nothing - when the operator or one of its
    // expressions is outside the subset, which ends the tail: the caller keeps running its own operators from there.
    bool add(const json::Node& op, const std::vector<std::string>& key_texts, const std::vector<std::string>& agg_texts);
    // Runs the tail over a group result: the rows FinalProject would send, as a JSON array (objects with sorted field
    // names, like Go's map marshalling).  Without an InitialProject the rows are the groups in n1gpu_result_to_json shape.
    std::string Run(const Result& r, i64* rows_out) const;
};

}  // namespace execution
}  // namespace n1
