// zignal-b200 :: lowering (symbolic tick) + host interpreter.  See zg_ir.hpp.
#include "zg_ir.hpp"

#include <algorithm>
#include <cstring>
#include <map>
#include <sstream>
#include <tuple>

namespace zg {

namespace {

constexpr int kBottom = -2;   // bottom_type (flowz.hpp:1004): current value of a fed-back wire
constexpr int kNoLine = -1;   // no_state    (flowz.hpp:128)

Dtype promote(Dtype a, Dtype b) { return (Dtype)std::max((int)a, (int)b); }

struct Builder {
    std::vector<IrNode> nodes;
    std::vector<IrLine> lines;
    std::map<std::tuple<int, int, int, int, uint64_t>, int> memo;
    bool cse = true;

    // Folding of literal-only sub-trees, e.g. -a2 or 2*pi in front of a wire.  Done in the node's
    // own dtype with the host's IEEE arithmetic, i.e. exactly what evaluating it every tick gives.
    bool fold(IrNode& n) const {
        auto cst = [&](int id) { return id >= 0 && nodes[id].op == IrOp::Const; };
        auto conv = [&](int id, Dtype to) {
            double v = nodes[id].value;   // stored exactly for every dtype
            return to == Dtype::F32 ? (double)(float)v : to == Dtype::I32 ? (double)(int32_t)v : v;
        };
        if (n.op == IrOp::Neg && cst(n.a)) {
            n = IrNode{IrOp::Const, n.dtype, -1, -1, -nodes[n.a].value};
            return true;
        }
        if ((n.op == IrOp::Add || n.op == IrOp::Sub || n.op == IrOp::Mul || n.op == IrOp::Div) &&
            cst(n.a) && cst(n.b) && n.dtype != Dtype::I32) {
            double a = conv(n.a, n.dtype), c = conv(n.b, n.dtype), r = 0;
            if (n.dtype == Dtype::F32) {
                float x = (float)a, y = (float)c, z = 0;
                switch (n.op) {
                    case IrOp::Add: z = x + y; break;
                    case IrOp::Sub: z = x - y; break;
                    case IrOp::Mul: z = x * y; break;
                    default: z = x / y; break;
                }
                r = (double)z;
            } else {
                switch (n.op) {
                    case IrOp::Add: r = a + c; break;
                    case IrOp::Sub: r = a - c; break;
                    case IrOp::Mul: r = a * c; break;
                    default: r = a / c; break;
                }
            }
            n = IrNode{IrOp::Const, n.dtype, -1, -1, r};
            return true;
        }
        return false;
    }

    int add(IrNode n) {
        if (cse) fold(n);
        uint64_t bits;
        std::memcpy(&bits, &n.value, 8);
        auto key = std::make_tuple((int)n.op, (int)n.dtype, n.a, n.b, bits);
        if (cse) {
            auto it = memo.find(key);
            if (it != memo.end()) return it->second;
        }
        nodes.push_back(n);
        int id = (int)nodes.size() - 1;
        if (cse) memo.emplace(key, id);
        return id;
    }
    int new_line(int depth) {
        IrLine l; l.depth = depth;
        lines.push_back(l);
        return (int)lines.size() - 1;
    }

    // The current value of a fed-back wire (the reference's bottom_type, flowz.hpp:1004) as a forward reference:
    // a Fwd node stands for it until binary_feedback has evaluated its promise part and binds it to the node that
    // is fed back.  Its dtype is a guess (lower() iterates until the guesses match what was bound).
    std::vector<int> fwd_nodes;                  // in creation order
    std::vector<int> fwd_target;                 // same order; -1 = not bound yet
    int add_fwd(Dtype guess) {
        nodes.push_back(IrNode{IrOp::Fwd, guess, (int)fwd_nodes.size(), -1, 0});
        fwd_nodes.push_back((int)nodes.size() - 1);
        fwd_target.push_back(-1);
        return (int)nodes.size() - 1;
    }
    void bind(int fwd_node, int target) { fwd_target[nodes[fwd_node].a] = target; }
};

using Vals = std::vector<int>;
using Lines = std::vector<int>;

template <class T>
std::vector<T> take(const std::vector<T>& v, int n) {        // tuple_take: N > size -> whole tuple
    if (n >= (int)v.size()) return v;
    return std::vector<T>(v.begin(), v.begin() + std::max(0, n));
}
template <class T>
std::vector<T> drop(const std::vector<T>& v, int n) {        // tuple_drop: N > size -> empty
    if (n >= (int)v.size()) return {};
    return std::vector<T>(v.begin() + std::max(0, n), v.end());
}
template <class T>
std::vector<T> cat(std::vector<T> a, const std::vector<T>& b) {
    a.insert(a.end(), b.begin(), b.end());
    return a;
}

struct Lowering {
    Builder& b;
    const std::vector<Dtype>& in_dtypes;
    const std::vector<Dtype>& fwd_guess;         // dtype of the k-th forward reference (F32 beyond its end)

    Lines make_node_state(const Expr& left, const Expr& right) {
        // StateCtor( tuple_take_( output_arity(_left), max_input_delays(_right) ) )  (:699, :707)
        std::vector<int> depths = take(max_input_delays(right), output_arity(left));
        Lines ls;
        for (int d : depths) ls.push_back(d > 0 ? b.new_line(d) : kNoLine);
        return ls;
    }

    void push_all(const Lines& node_state, const Vals& vals, const char* who) {
        // tuple_for_each over the shorter of the two (tuple_tools.hpp:196-221)
        size_t m = std::min(node_state.size(), vals.size());
        for (size_t i = 0; i < m; ++i) {
            if (node_state[i] == kNoLine) continue;          // rotate_push_back(no_state, y): no-op
            if (vals[i] == kBottom)
                throw Error(std::string(who) + ": a fed-back wire is pushed before it has a value");
            b.lines[node_state[i]].src = vals[i];
        }
    }

    Vals eval(const Expr& e, const Vals& input, const Lines& in_state) {
        switch (e.op) {
            case Op::Placeholder: {                          // place_the_holder :941-948
                if (e.k > (int)input.size())
                    throw Error("placeholder _" + std::to_string(e.k) + " reads past the " +
                                std::to_string(input.size()) + " wires available at this point");
                return {input[e.k - 1]};
            }
            case Op::Delay: {                                // place_delay :950-958
                if (e.k > (int)in_state.size() || in_state[e.k - 1] == kNoLine)
                    throw Error("_" + std::to_string(e.k) + "[_" + std::to_string(e.n) +
                                "]: this wire has no delay line here (no_state, flowz.hpp:1059/1148)");
                int line = in_state[e.k - 1];
                if (e.n > b.lines[line].depth)
                    // binary_feedback hands its future part ALL external inputs (:1043-1047, drop<min(0, ...)> and its
                    // TODO) while the delay analysis that sizes the state (:443-506) gives it the wires after the
                    // promise part's; where they disagree the reference indexes before the start of a std::array
                    throw Error("_" + std::to_string(e.k) + "[_" + std::to_string(e.n) + "] reaches past the delay line the " +
                                "reference allocates for this wire (" + std::to_string(b.lines[line].depth) +
                                " deep): out-of-bounds in the reference (flowz.hpp:950-958, 1043-1047)");
                IrNode n{IrOp::DRead, Dtype::F32, line, e.n, 0};
                return {b.add(n)};
            }
            case Op::Const:
                if (e.dtype == Dtype::C64 || e.dtype == Dtype::C128)
                    throw Error("complex terminals are typed (zg_expr_result_types) but not evaluated: the reference's "
                                "compile() keeps float state only (flowz.hpp:1245)");
                return {b.add(IrNode{IrOp::Const, e.dtype, -1, -1, e.value})};
            case Op::Param: return {b.add(IrNode{IrOp::Param, Dtype::F32, e.k, -1, 0})};
            case Op::Neg: {
                int a = scalar(e, *e.ch[0], input, in_state);
                return {b.add(IrNode{IrOp::Neg, b.nodes[a].dtype, a, -1, 0})};
            }
            case Op::Add: case Op::Sub: case Op::Mul: case Op::Div: {
                int a = scalar(e, *e.ch[0], input, in_state);
                int c = scalar(e, *e.ch[1], input, in_state);
                IrOp op = e.op == Op::Add ? IrOp::Add : e.op == Op::Sub ? IrOp::Sub
                        : e.op == Op::Mul ? IrOp::Mul : IrOp::Div;
                return {b.add(IrNode{op, promote(b.nodes[a].dtype, b.nodes[c].dtype), a, c, 0})};
            }
            case Op::Chan: {                                 // make_tuple(eval(l), eval(r)), same env :765-768
                stateless(*e.ch[0], "channel (a , b)");
                stateless(*e.ch[1], "channel (a , b)");
                return cat(eval(*e.ch[0], input, in_state), eval(*e.ch[1], input, in_state));
            }
            case Op::Par: {                                  // parallel :1076-1101
                int inL = input_arity(*e.ch[0]);
                Vals l = eval(*e.ch[0], take(input, inL), take(in_state, inL));
                Vals r = eval(*e.ch[1], drop(input, inL), drop(in_state, inL));
                return cat(l, r);
            }
            case Op::Seq: {                                  // sequence :960-1001
                const Expr& L = *e.ch[0];
                const Expr& R = *e.ch[1];
                int inL = input_arity(L), inR = input_arity(R);
                Lines node_state = make_node_state(L, R);
                Vals left_result = eval(L, take(input, inL), take(in_state, inL));
                Vals right_input = cat(left_result, drop(input, inL));
                Lines right_delayed = cat(node_state, drop(in_state, inL));
                Vals right_result = eval(R, right_input, right_delayed);
                push_all(node_state, left_result, "sequence");
                Vals out = cat(right_result, drop(left_result, inR));
                return cat(out, drop(input, inL + (int)left_result.size()));
            }
            case Op::Bfb: {                                  // binary_feedback :1031-1074
                const Expr& L = *e.ch[0];   // promise
                const Expr& R = *e.ch[1];   // future
                int inL = input_arity(L), outL = output_arity(L), outR = output_arity(R);
                if (inL < outR)
                    throw Error("binary_feedback: promise part takes fewer inputs than the future part "
                                "produces (ill-formed in the reference too, flowz.hpp:1046)");
                Lines node_state = make_node_state(L, R);
                Vals future_input = cat(Vals(outL, kBottom), input);               // :1043-1047 (drop<0>)
                Lines future_delayed = cat(node_state, in_state);                  // :1048-1050
                Vals result = eval(R, future_input, future_delayed);
                Vals promise_input = cat(result, take(input, inL - outR));         // :1057
                Lines promise_delayed = cat(Lines(outL, kNoLine), take(in_state, inL - outR));
                Vals promise_result = eval(L, promise_input, promise_delayed);
                push_all(node_state, promise_result, "binary_feedback");
                return result;                                                     // :1069
            }
            case Op::Fb: {
                // A feedback kept whole (zg_expr.cpp: the reference cannot split it, or its split would touch the
                // current value of a fed-back wire): the first out(x) inputs of x are x's own outputs.  They enter as
                // forward references, bound once x has been walked; delayed reads of them come from lines this node
                // owns, as in binary_feedback.  Only a loop without a delay is an error (resolve_forward_references).
                const Expr& X = *e.ch[0];
                const int n = output_arity(X);
                std::vector<int> depths = take(max_input_delays(X), n);
                Lines node_state;
                for (int d : depths) node_state.push_back(d > 0 ? b.new_line(d) : kNoLine);
                node_state.resize(n, kNoLine);
                Vals fed(n);
                for (int i = 0; i < n; ++i) {
                    const size_t k = b.fwd_nodes.size();
                    fed[i] = b.add_fwd(k < fwd_guess.size() ? fwd_guess[k] : Dtype::F32);
                }
                Vals result = eval(X, cat(fed, input), cat(node_state, in_state));
                if ((int)result.size() < n) throw Error("feedback: the expression returns fewer wires than it feeds back");
                push_all(node_state, result, "feedback");
                for (int i = 0; i < n; ++i) b.bind(fed[i], result[i]);
                return result;
            }
        }
        throw Error("unknown expression node");
    }

    // operand of an arithmetic node: must be a single, real value
    int scalar(const Expr& parent, const Expr& child, const Vals& input, const Lines& in_state) {
        stateless(child, "arithmetic");
        Vals v = eval(child, input, in_state);
        if (v.size() != 1)
            throw Error("arithmetic on a multi-wire sub-expression: " + to_string(parent));
        if (v[0] == kBottom)
            throw Error("a fed-back wire is used undelayed inside its own loop: " + to_string(parent));
        return v[0];
    }

    // build_state gives no state to children of arithmetic / channel nodes (flowz.hpp:719), so a
    // stateful combinator below them does not compile in the reference.
    static void stateless(const Expr& e, const char* where) {
        if (e.op == Op::Seq || e.op == Op::Par || e.op == Op::Bfb || e.op == Op::Fb)
            throw Error(std::string("combinator below ") + where +
                        " node is not supported (reference: build_state yields no state there, "
                        "flowz.hpp:719): " + to_string(e));
        for (auto& c : e.ch) stateless(*c, where);
    }
};

// Forward references (Builder::add_fwd) -> the nodes they were bound to; then a topological order, because a node that
// used a fed-back wire directly was created before the node that feeds it.  A cycle here is a loop without a delay --
// delayed reads (DRead) depend on nothing within a tick.  Programs without feedback are left untouched, programs whose
// feedbacks were split canonically keep their order (their forward references are never read).
// Returns false when a dtype guess was wrong (fwd_guess is updated; the caller lowers again).
bool resolve_forward_references(Builder& b, Vals& outs, std::vector<Dtype>& fwd_guess) {
    const size_t nf = b.fwd_nodes.size();
    if (nf == 0) return true;
    auto resolve = [&](int id) {
        size_t hops = 0;
        while (id >= 0 && b.nodes[id].op == IrOp::Fwd) {
            id = b.fwd_target[b.nodes[id].a];
            if (id < 0) throw Error("a fed-back wire was never given a value (internal)");
            if (++hops > nf) throw Error("feedback loop without a delay: a fed-back wire is fed by itself");
        }
        return id;
    };
    bool settled = true;
    fwd_guess.resize(std::max(fwd_guess.size(), nf), Dtype::F32);
    for (size_t k = 0; k < nf; ++k) {
        const Dtype actual = b.nodes[resolve(b.fwd_nodes[k])].dtype;
        if (b.nodes[b.fwd_nodes[k]].dtype != actual) { fwd_guess[k] = actual; settled = false; }
    }
    if (!settled) return false;

    auto binary = [](IrOp op) { return op == IrOp::Add || op == IrOp::Sub || op == IrOp::Mul || op == IrOp::Div; };
    bool used = false;                                       // is any forward reference actually read?
    for (IrNode& n : b.nodes) {
        if (n.op == IrOp::Neg || binary(n.op)) {
            const int a = resolve(n.a);
            used = used || a != n.a;
            n.a = a;
        }
        if (binary(n.op)) {
            const int c = resolve(n.b);
            used = used || c != n.b;
            n.b = c;
        }
    }
    for (int& o : outs) { const int r = resolve(o); used = used || r != o; o = r; }
    for (IrLine& l : b.lines) if (l.src >= 0) { const int r = resolve(l.src); used = used || r != l.src; l.src = r; }
    if (!used) return true;

    // depth-first post-order over the nodes in index order: the identity wherever the order already was topological
    const int n = (int)b.nodes.size();
    std::vector<int> order, where(n, -1);
    std::vector<char> mark(n, 0);                            // 1 = on the stack, 2 = emitted
    order.reserve(n);
    std::vector<std::pair<int, int>> stack;
    for (int root = 0; root < n; ++root) {
        if (mark[root]) continue;
        stack.push_back({root, 0});
        mark[root] = 1;
        while (!stack.empty()) {
            auto& [id, next] = stack.back();
            const IrNode& nd = b.nodes[id];
            const int deps[2] = {(nd.op == IrOp::Neg || binary(nd.op)) ? nd.a : -1, binary(nd.op) ? nd.b : -1};
            if (next < 2) {
                const int d = deps[next++];
                if (d < 0 || mark[d] == 2) continue;
                if (mark[d] == 1)
                    throw Error("feedback loop without a delay: a fed-back wire is used undelayed by the expression that "
                                "produces it (put a _k[_n] somewhere in the loop)");
                mark[d] = 1;
                stack.push_back({d, 0});
            } else {
                mark[id] = 2;
                where[id] = (int)order.size();
                order.push_back(id);
                stack.pop_back();
            }
        }
    }
    std::vector<IrNode> sorted;
    sorted.reserve(n);
    for (int id : order) {
        IrNode nd = b.nodes[id];
        if (nd.op == IrOp::Neg || binary(nd.op)) nd.a = where[nd.a];
        if (binary(nd.op)) nd.b = where[nd.b];
        sorted.push_back(nd);
    }
    b.nodes.swap(sorted);
    for (int& o : outs) o = where[o];
    for (IrLine& l : b.lines) if (l.src >= 0) l.src = where[l.src];
    for (int& f : b.fwd_nodes) f = where[f];
    b.memo.clear();                                          // (node ids changed; the rebuild below re-runs CSE)
    return true;
}

}  // namespace

bool Ir::all_f32() const {
    for (auto& n : nodes) if (n.dtype != Dtype::F32) return false;
    for (auto d : in_dtypes) if (d != Dtype::F32) return false;
    return true;
}

std::string Ir::dump() const {
    static const char* opn[] = {"in", "const", "param", "dread", "neg", "add", "sub", "mul", "div", "fwd"};
    static const char* dtn[] = {"i32", "f32", "f64"};
    std::ostringstream os;
    os << "graph n_in=" << n_in << " n_out=" << n_out << " n_params=" << n_params
       << " n_state=" << n_state << "\n";
    for (size_t i = 0; i < lines.size(); ++i)
        os << "  line" << i << " depth=" << lines[i].depth << " offset=" << lines[i].offset
           << " <- %" << lines[i].src << "\n";
    for (size_t i = 0; i < nodes.size(); ++i) {
        const IrNode& n = nodes[i];
        os << "  %" << i << " = " << opn[(int)n.op] << "." << dtn[(int)n.dtype];
        switch (n.op) {
            case IrOp::In: os << " " << n.a; break;
            case IrOp::Param: os << " $" << n.a; break;
            case IrOp::Const: { char buf[48]; std::snprintf(buf, sizeof buf, " %.9g", n.value); os << buf; break; }
            case IrOp::DRead: os << " line" << n.a << "[-" << n.b << "]"; break;
            case IrOp::Neg: os << " %" << n.a; break;
            default: os << " %" << n.a << ", %" << n.b; break;
        }
        os << "\n";
    }
    os << "  out";
    for (int o : outs) os << " %" << o;
    os << "\n";
    return os.str();
}

Ir lower(const Expr& canonical, const std::vector<Dtype>& in_dtypes, const LowerOptions& opt) {
    // The callable takes as many arguments as the USER's expression has inputs (arity_t, flowz.hpp:1238), which is what
    // in_dtypes has.  The canonical tree can count more -- the split of a feedback can move a sub-expression with
    // unused inputs into the promise part -- and nothing is wrong as long as no placeholder actually reads past the
    // wires that exist (Lowering::eval checks that where it happens).
    const int n_in = (int)in_dtypes.size();

    Builder b;
    Vals outs;
    std::vector<Dtype> fwd_guess;                            // dtypes of the forward references, by creation order
    for (int pass = 0;; ++pass) {
        b = Builder();
        b.cse = opt.cse;
        Vals input;
        for (int i = 0; i < n_in; ++i) input.push_back(b.add(IrNode{IrOp::In, in_dtypes[i], i, -1, 0}));
        Lowering lw{b, in_dtypes, fwd_guess};
        outs = lw.eval(canonical, input, Lines{});           // in_state = std::tuple<>{} (:1196)
        for (int o : outs)
            if (o == kBottom) throw Error("graph output is an unresolved fed-back wire");
        if (resolve_forward_references(b, outs, fwd_guess)) break;
        if (pass >= 4) throw Error("the types of the fed-back wires do not settle (internal)");
    }

    // ---- line merging: lines fed by the same node hold the same history --------------------
    std::vector<int> line_map(b.lines.size());
    std::vector<IrLine> lines;
    {
        std::map<int, int> by_src;
        for (size_t i = 0; i < b.lines.size(); ++i) {
            const IrLine& l = b.lines[i];
            if (l.src < 0) throw Error("delay line without a writer (internal)");
            auto it = opt.merge_lines ? by_src.find(l.src) : by_src.end();
            if (it != by_src.end()) {
                line_map[i] = it->second;
                lines[it->second].depth = std::max(lines[it->second].depth, l.depth);
            } else {
                line_map[i] = (int)lines.size();
                if (opt.merge_lines) by_src[l.src] = (int)lines.size();
                lines.push_back(l);
            }
        }
    }

    // ---- rebuild: remap DRead lines, re-run CSE, drop dead nodes and unread lines ----------
    std::vector<char> live(b.nodes.size(), 0);
    {
        // a line is needed only if something live reads it; iterate to a fixed point
        std::vector<char> line_live(lines.size(), 0);
        std::vector<int> work(outs.begin(), outs.end());
        auto mark = [&](int id) { if (id >= 0 && !live[id]) { live[id] = 1; work.push_back(id); } };
        for (int o : outs) live[o] = 1;
        while (!work.empty()) {
            int id = work.back(); work.pop_back();
            const IrNode& n = b.nodes[id];
            if (n.op == IrOp::DRead) {
                int l = line_map[n.a];
                if (!line_live[l]) { line_live[l] = 1; mark(lines[l].src); }
            } else if (n.op == IrOp::Neg) {
                mark(n.a);
            } else if (n.op == IrOp::Add || n.op == IrOp::Sub || n.op == IrOp::Mul || n.op == IrOp::Div) {
                mark(n.a); mark(n.b);
            }
        }
        std::vector<int> compact(lines.size(), -1);
        std::vector<IrLine> kept;
        for (size_t l = 0; l < lines.size(); ++l)
            if (line_live[l]) { compact[l] = (int)kept.size(); kept.push_back(lines[l]); }
        for (auto& m : line_map) m = compact[m];
        lines.swap(kept);
    }

    Builder c;
    c.cse = opt.cse;
    std::vector<int> remap(b.nodes.size(), -1);
    for (size_t i = 0; i < b.nodes.size(); ++i) {
        IrNode n = b.nodes[i];
        if (n.op == IrOp::In) { remap[i] = c.add(n); continue; }   // inputs keep their slots
        if (!live[i]) continue;
        switch (n.op) {
            case IrOp::DRead: n.a = line_map[n.a]; break;
            case IrOp::Neg: n.a = remap[n.a]; break;
            case IrOp::Add: case IrOp::Sub: case IrOp::Mul: case IrOp::Div:
                n.a = remap[n.a]; n.b = remap[n.b]; break;
            default: break;
        }
        remap[i] = c.add(n);
    }

    Ir ir;
    ir.n_in = n_in;
    ir.in_dtypes = in_dtypes;
    ir.n_params = n_params(canonical);
    ir.nodes = std::move(c.nodes);
    for (int o : outs) ir.outs.push_back(remap[o]);
    ir.n_out = (int)ir.outs.size();
    int64_t off = 0;
    for (auto& l : lines) {
        l.src = remap[l.src];
        l.offset = (int)off;
        off += l.depth;
        if (off > (int64_t(1) << 26)) throw Error("graph keeps more than 2^26 floats of delay-line state per voice");
    }
    ir.n_state = (int)off;
    ir.lines = std::move(lines);
    return ir;
}

// ------------------------------------------------------------------------------------------------
// host interpreter
// ------------------------------------------------------------------------------------------------

namespace {

// double -> C++ int at the ABI: out-of-range and NaN values (undefined behaviour as a plain cast) saturate / become 0
inline int32_t to_i32(double v) {
    if (!(v == v)) return 0;
    return v >= 2147483647.0 ? INT32_MAX : v <= -2147483648.0 ? INT32_MIN : (int32_t)v;
}

union Cell { int32_t i; float f; double d; };

inline double as_f64(Cell c, Dtype t) { return t == Dtype::I32 ? (double)c.i : t == Dtype::F32 ? (double)c.f : c.d; }
inline float as_f32(Cell c, Dtype t) { return t == Dtype::I32 ? (float)c.i : t == Dtype::F32 ? c.f : (float)c.d; }

template <class F32, class F64, class I32>
inline Cell arith(Dtype t, Cell a, Dtype ta, Cell b, Dtype tb, F32 f32, F64 f64, I32 i32) {
    Cell r;
    switch (t) {
        case Dtype::I32: r.i = i32(a.i, b.i); break;
        case Dtype::F32: r.f = f32(as_f32(a, ta), as_f32(b, tb)); break;
        case Dtype::F64: r.d = f64(as_f64(a, ta), as_f64(b, tb)); break;
        default: break;                                  // complex: never lowered (Builder::eval rejects it)
    }
    return r;
}

inline void run_tick(const Ir& ir, float* state, const float* params, Cell* v) {
    const size_t n = ir.nodes.size();
    for (size_t i = 0; i < n; ++i) {
        const IrNode& nd = ir.nodes[i];
        switch (nd.op) {
            case IrOp::In: break;  // pre-filled
            case IrOp::Const:
                if (nd.dtype == Dtype::I32) v[i].i = to_i32(nd.value);
                else if (nd.dtype == Dtype::F32) v[i].f = (float)nd.value;
                else v[i].d = nd.value;
                break;
            case IrOp::Param: v[i].f = params[nd.a]; break;
            case IrOp::DRead: {
                const IrLine& l = ir.lines[nd.a];
                v[i].f = state[l.offset + l.depth - nd.b];
                break;
            }
            case IrOp::Neg:
                if (nd.dtype == Dtype::I32) v[i].i = (int32_t)(0u - (uint32_t)v[nd.a].i);   // wraps like the other int ops
                else if (nd.dtype == Dtype::F32) v[i].f = -v[nd.a].f;
                else v[i].d = -v[nd.a].d;
                break;
            case IrOp::Add:
                v[i] = arith(nd.dtype, v[nd.a], ir.nodes[nd.a].dtype, v[nd.b], ir.nodes[nd.b].dtype,
                             [](float a, float b) { return a + b; }, [](double a, double b) { return a + b; },
                             [](int32_t a, int32_t b) { return (int32_t)((uint32_t)a + (uint32_t)b); });
                break;
            case IrOp::Sub:
                v[i] = arith(nd.dtype, v[nd.a], ir.nodes[nd.a].dtype, v[nd.b], ir.nodes[nd.b].dtype,
                             [](float a, float b) { return a - b; }, [](double a, double b) { return a - b; },
                             [](int32_t a, int32_t b) { return (int32_t)((uint32_t)a - (uint32_t)b); });
                break;
            case IrOp::Mul:
                v[i] = arith(nd.dtype, v[nd.a], ir.nodes[nd.a].dtype, v[nd.b], ir.nodes[nd.b].dtype,
                             [](float a, float b) { return a * b; }, [](double a, double b) { return a * b; },
                             [](int32_t a, int32_t b) { return (int32_t)((int64_t)a * (int64_t)b); });
                break;
            case IrOp::Div:
                v[i] = arith(nd.dtype, v[nd.a], ir.nodes[nd.a].dtype, v[nd.b], ir.nodes[nd.b].dtype,
                             [](float a, float b) { return a / b; }, [](double a, double b) { return a / b; },
                             [](int32_t a, int32_t b) { return b == 0 ? 0 : b == -1 ? (int32_t)(0u - (uint32_t)a) : a / b; });
                break;
            case IrOp::Fwd: break;                       // never survives lower()
        }
    }
    // rotate_push_back (:130-148) for every line, after everything has been read
    for (const IrLine& l : ir.lines) {
        float* s = state + l.offset;
        float y = as_f32(v[l.src], ir.nodes[l.src].dtype);
        for (int j = 1; j < l.depth; ++j) s[j - 1] = s[j];
        s[l.depth - 1] = y;
    }
}

}  // namespace

void host_tick(const Ir& ir, float* state, const float* params, const double* in, double* out) {
    std::vector<Cell> v(ir.nodes.size());
    for (int i = 0; i < ir.n_in; ++i) {
        switch (ir.in_dtypes[i]) {
            case Dtype::I32: v[i].i = to_i32(in[i]); break;
            case Dtype::F32: v[i].f = (float)in[i]; break;
            case Dtype::F64: v[i].d = in[i]; break;
            default: break;
        }
    }
    run_tick(ir, state, params, v.data());
    for (int o = 0; o < ir.n_out; ++o) out[o] = as_f64(v[ir.outs[o]], ir.nodes[ir.outs[o]].dtype);
}

void host_block_f32(const Ir& ir, float* state, const float* params, const float* const* in,
                    float* const* out, long n_samples, long in_stride, long out_stride) {
    std::vector<Cell> v(ir.nodes.size());
    for (long t = 0; t < n_samples; ++t) {
        for (int i = 0; i < ir.n_in; ++i) {
            float x = in[i][t * in_stride];
            switch (ir.in_dtypes[i]) {
                case Dtype::I32: v[i].i = to_i32(x); break;
                case Dtype::F32: v[i].f = x; break;
                case Dtype::F64: v[i].d = x; break;
                default: break;
            }
        }
        run_tick(ir, state, params, v.data());
        for (int o = 0; o < ir.n_out; ++o)
            out[o][t * out_stride] = as_f32(v[ir.outs[o]], ir.nodes[ir.outs[o]].dtype);
    }
}

// ---- long delay lines: kernel-side program ------------------------------------------------------------------------

Ir split_long_lines(const Ir& ir, int reg_depth, int far, RingPlan& rp) {
    Ir k = ir;
    rp = RingPlan{};
    const int L = (int)ir.lines.size();
    std::vector<char> is_long(L, 0);
    std::vector<int> near_depth(L, 0);         // deepest near read of a long line
    for (int l = 0; l < L; ++l) is_long[l] = ir.lines[l].depth > reg_depth;
    for (const IrNode& n : ir.nodes)
        if (n.op == IrOp::DRead && is_long[n.a] && n.b < far) near_depth[n.a] = std::max(near_depth[n.a], n.b);
    // far reads -> extra inputs (one per distinct (line, n))
    std::vector<std::pair<int, int>> seen;
    for (IrNode& n : k.nodes) {
        if (n.op != IrOp::DRead || !is_long[n.a] || n.b < far) continue;
        int idx = -1;
        for (size_t i = 0; i < seen.size(); ++i)
            if (seen[i].first == n.a && seen[i].second == n.b) idx = (int)i;
        if (idx < 0) {
            idx = (int)seen.size();
            seen.push_back({n.a, n.b});
            rp.taps.push_back({n.a, n.b});
        }
        n.op = IrOp::In;
        n.a = ir.n_in + idx;
        n.b = -1;
    }
    k.n_in = ir.n_in + (int)rp.taps.size();
    k.in_dtypes.resize(k.n_in, Dtype::F32);
    // pushed values of long lines -> extra outputs
    for (int l = 0; l < L; ++l)
        if (is_long[l]) {
            rp.out_lines.push_back(l);
            k.outs.push_back(ir.lines[l].src);
        }
    k.n_out = (int)k.outs.size();
    // kernel lines: short lines as they are, long lines shrunk to their register window (possibly empty)
    k.lines.clear();
    std::vector<int> new_index(L, -1);
    int off = 0;
    for (int l = 0; l < L; ++l) {
        const IrLine& o = ir.lines[l];
        const int depth = is_long[l] ? near_depth[l] : o.depth;
        if (depth == 0) continue;
        IrLine nl;
        nl.depth = depth;
        nl.offset = off;
        nl.src = o.src;
        new_index[l] = (int)k.lines.size();
        k.lines.push_back(nl);
        for (int j = 0; j < depth; ++j) {
            KernelSlot s;
            if (is_long[l]) {                  // window slot j holds the value pushed depth - j ticks ago
                s.row = o.offset;
                s.ring_depth = o.depth;
                s.ago = depth - j;
            } else {
                s.row = o.offset + j;
            }
            rp.slots.push_back(s);
        }
        off += depth;
    }
    k.n_state = off;
    for (IrNode& n : k.nodes)
        if (n.op == IrOp::DRead) n.a = new_index[n.a];
    return k;
}

}  // namespace zg
