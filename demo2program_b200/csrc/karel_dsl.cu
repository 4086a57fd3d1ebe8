// Karel DSL: shift-reduce parser, interpreter and program-level evaluation metrics (host CPU).
//
// Restates, as one AST interpreter, what the reference does with Python closures:
//   * parser      karel_env/dsl/dsl_parse.py:4-13 (check_and_apply), 24-262 (rules, in order), 252-265 (parse)
//   * semantics   the rule closures, incl. the call counter `n` and MAX_FUNC_CALL = 100 time-out
//   * world       karel_env/karel.py:33-185 (Karel_world: perception primitives, state_transition)
//   * metrics     models/model_full.py:602-616 (check_correct_syntax), 747-787 (generate_program_output_karel),
//                 870-897 (CompareDemoAndExecution), 712-727 + karel_env/dsl/dsl_enum_program.py (exact compare)
// Token ids are the reference vocabulary order (dsl_prob.py:13-28 with INT expanded).
//
// This file has no device code; it lives in libd2p.so so that the evaluation path has one native
// library.  The reference interprets B*(k+test_k)*2 programs per evaluation step in Python.
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#include "common.cuh"

namespace d2p {
namespace karel {

// ---- vocabulary -------------------------------------------------------------------------------
enum Tok : int {
    T_DEF = 0, T_RUN = 1, T_MOPEN = 2, T_MCLOSE = 3, T_MOVE = 4, T_TURNRIGHT = 5, T_TURNLEFT = 6,
    T_PICK = 7, T_PUT = 8, T_ROPEN = 9, T_RCLOSE = 10, T_INT0 = 11, T_INT19 = 30, T_REPEAT = 31,
    T_COPEN = 32, T_CCLOSE = 33, T_IOPEN = 34, T_ICLOSE = 35, T_EOPEN = 36, T_ECLOSE = 37, T_IF = 38,
    T_IFELSE = 39, T_ELSE = 40, T_FRONT = 41, T_LEFT = 42, T_RIGHT = 43, T_MARKERS = 44, T_NOMARKERS = 45,
    T_NOT = 46, T_WOPEN = 47, T_WCLOSE = 48, T_WHILE = 49, N_TOK = 50
};
// grammar symbols: terminals are their token ids, non-terminals follow
enum Sym : int {
    S_PROG = 100, S_STMT, S_STMT_STMT, S_WHILE, S_REPEAT, S_ACTION, S_IF, S_IFELSE, S_COND, S_COND_WN, S_CSTE
};
constexpr int MAX_FUNC_CALL = 100;   // dsl_parse.py:21
constexpr int DEPTH = 16, MAX_NUM_MARKER = 10;

// ---- AST --------------------------------------------------------------------------------------
enum Kind { K_PROG, K_STMT, K_SEQ, K_IF, K_IFELSE, K_WHILE, K_REPEAT, K_COND_WRAP, K_COND_NOT, K_COND_PRIM,
            K_ACTION, K_CSTE };
struct Node {
    Kind kind;
    int arg = 0;                      // action index / perception index / repeat count
    Node* a = nullptr;                // cond or first statement
    Node* b = nullptr;                // statement / second statement
    Node* c = nullptr;                // else statement
};
struct Arena {
    std::vector<std::unique_ptr<Node>> nodes;
    Node* make(Kind k, int arg = 0, Node* a = nullptr, Node* b = nullptr, Node* c = nullptr) {
        nodes.emplace_back(new Node{k, arg, a, b, c});
        return nodes.back().get();
    }
};

// ---- parser -----------------------------------------------------------------------------------
struct Item { int sym; Node* node; };

static bool tail_is(const std::vector<Item>& q, std::initializer_list<int> pat) {
    if (q.size() < pat.size()) return false;
    size_t off = q.size() - pat.size(), i = 0;
    for (int s : pat) if (q[off + i++].sym != s) return false;
    return true;
}
static void reduce(std::vector<Item>& q, size_t n, int sym, Node* node) {
    q.resize(q.size() - n);
    q.push_back({sym, node});
}

// tries the rules in the reference's order; applies the first whose right-hand side matches the
// top of the stack
static bool apply_one_rule(std::vector<Item>& q, Arena& ar) {
    auto at = [&](size_t from_end) { return q[q.size() - from_end].node; };
    if (tail_is(q, {T_DEF, T_RUN, T_MOPEN, S_STMT, T_MCLOSE})) {
        reduce(q, 5, S_PROG, ar.make(K_PROG, 0, nullptr, at(2)));
        return true;
    }
    for (int s : {S_WHILE, S_REPEAT, S_STMT_STMT, S_ACTION, S_IF, S_IFELSE})
        if (tail_is(q, {s})) { reduce(q, 1, S_STMT, ar.make(K_STMT, 0, nullptr, at(1))); return true; }
    if (tail_is(q, {S_STMT, S_STMT})) {
        reduce(q, 2, S_STMT_STMT, ar.make(K_SEQ, 0, at(2), at(1)));
        return true;
    }
    if (tail_is(q, {T_IF, T_COPEN, S_COND, T_CCLOSE, T_IOPEN, S_STMT, T_ICLOSE})) {
        reduce(q, 7, S_IF, ar.make(K_IF, 0, at(5), at(2)));
        return true;
    }
    if (tail_is(q, {T_IFELSE, T_COPEN, S_COND, T_CCLOSE, T_IOPEN, S_STMT, T_ICLOSE, T_ELSE, T_EOPEN, S_STMT,
                    T_ECLOSE})) {
        reduce(q, 11, S_IFELSE, ar.make(K_IFELSE, 0, at(9), at(6), at(2)));
        return true;
    }
    if (tail_is(q, {T_WHILE, T_COPEN, S_COND, T_CCLOSE, T_WOPEN, S_STMT, T_WCLOSE})) {
        reduce(q, 7, S_WHILE, ar.make(K_WHILE, 0, at(5), at(2)));
        return true;
    }
    if (tail_is(q, {T_REPEAT, S_CSTE, T_ROPEN, S_STMT, T_RCLOSE})) {
        reduce(q, 5, S_REPEAT, ar.make(K_REPEAT, at(4)->arg, nullptr, at(2)));
        return true;
    }
    if (tail_is(q, {S_COND_WN})) { reduce(q, 1, S_COND, ar.make(K_COND_WRAP, 0, at(1))); return true; }
    if (tail_is(q, {T_NOT, T_COPEN, S_COND, T_CCLOSE})) {
        reduce(q, 4, S_COND, ar.make(K_COND_NOT, 0, at(2)));
        return true;
    }
    static const int prims[5] = {T_FRONT, T_LEFT, T_RIGHT, T_MARKERS, T_NOMARKERS};
    for (int i = 0; i < 5; ++i)
        if (tail_is(q, {prims[i]})) { reduce(q, 1, S_COND_WN, ar.make(K_COND_PRIM, i)); return true; }
    // action indices of karel.py:24-30: 0 move, 1 turn left, 2 turn right, 3 pick, 4 put
    static const int acts[5] = {T_MOVE, T_TURNLEFT, T_TURNRIGHT, T_PICK, T_PUT};
    for (int i = 0; i < 5; ++i)
        if (tail_is(q, {acts[i]})) { reduce(q, 1, S_ACTION, ar.make(K_ACTION, i)); return true; }
    if (!q.empty() && q.back().sym >= T_INT0 && q.back().sym <= T_INT19) {
        int v = q.back().sym - T_INT0;
        reduce(q, 1, S_CSTE, ar.make(K_CSTE, v));
        return true;
    }
    return false;
}

// dsl_parse.py:252-265.  Like the reference, the loop ends as soon as the input is consumed and ONE
// symbol is left, whatever that symbol is (a lone `move` "parses").  Returns nullptr on failure.
static Node* parse(const int* tokens, int len, Arena& ar, int* root_sym) {
    if (len <= 0) return nullptr;                 // the reference raises IndexError on an empty string
    for (int i = 0; i < len; ++i) if (tokens[i] < 0 || tokens[i] >= N_TOK) return nullptr;
    std::vector<Item> q;
    int pos = 0;
    bool applied = false;
    while (pos < len || q.size() != 1) {
        if (applied) applied = false;
        else q.push_back({tokens[pos++], nullptr});
        applied = apply_one_rule(q, ar);
        if (!applied && pos >= len) return nullptr;
    }
    if (root_sym) *root_sym = q[0].sym;
    return q[0].node;
}

// ---- world (karel.py) -------------------------------------------------------------------------
struct World {
    int h, w;
    bool make_error;
    std::vector<uint8_t> s;                   // [h][w][16], 0/1
    std::vector<std::vector<uint8_t>> hist;   // s_h
    bool broken = false;                      // no hero in the state: the reference would raise
    uint8_t& at(int x, int y, int z) { return s[((size_t)x * w + y) * DEPTH + z]; }
    bool locate(int* x, int* y, int* z) {     // get_location: first hit in (x, y, z) order
        for (int i = 0; i < h; ++i)
            for (int j = 0; j < w; ++j)
                for (int d = 0; d < 4; ++d)
                    if (at(i, j, d)) { *x = i; *y = j; *z = d; return true; }
        broken = true;
        return false;
    }
    bool neighbor_clear(int face) {           // 0 front, 1 left, 2 right (karel.py:67-105)
        static const int dx[3][4] = {{-1, 0, 1, 0}, {0, -1, 0, 1}, {0, 1, 0, -1}};
        static const int dy[3][4] = {{0, 1, 0, -1}, {-1, 0, 1, 0}, {1, 0, -1, 0}};
        int x, y, z;
        if (!locate(&x, &y, &z)) return false;
        int nx = x + dx[face][z], ny = y + dy[face][z];
        if (nx >= h || nx < 0 || ny >= w || ny < 0) return false;
        return !at(nx, ny, 4);
    }
    int markers_here() {                      // sum(s[x, y, 6:])
        int x, y, z, n = 0;
        if (!locate(&x, &y, &z)) return 0;
        for (int d = 6; d < DEPTH; ++d) n += at(x, y, d) ? 1 : 0;
        return n;
    }
    bool perceive(int p) {
        switch (p) {
            case 0: return neighbor_clear(0);
            case 1: return neighbor_clear(1);
            case 2: return neighbor_clear(2);
            case 3: return markers_here() > 0;
            default: return markers_here() == 0;
        }
    }
    // state_transition (karel.py:138-185); false = the reference raises RuntimeError
    bool act(int a) {
        int x, y, z;
        if (!locate(&x, &y, &z)) return false;
        if (a == 0) {
            if (neighbor_clear(0)) {
                static const int fx[4] = {-1, 0, 1, 0}, fy[4] = {0, 1, 0, -1};
                int nx = x + fx[z], ny = y + fy[z];
                for (int d = 0; d < 4; ++d) { at(nx, ny, d) = at(x, y, d); }
                for (int d = 0; d < 4; ++d) at(x, y, d) = 0;
            } else {
                if (make_error) return false;
                for (int d = 0; d < 4; ++d) at(x, y, d) = 0;
                at(x, y, (z + 2) % 4) = 1;                      // turn 180
            }
        } else if (a == 1 || a == 2) {
            for (int d = 0; d < 4; ++d) at(x, y, d) = 0;
            at(x, y, ((a * 2 - 3 + z) % 4 + 4) % 4) = 1;
        } else if (a == 3 || a == 4) {
            int num = 0;                                         // argmax(s[x, y, 5:])
            for (int d = 5; d < DEPTH; ++d) if (at(x, y, d)) { num = d - 5; break; }
            int nn = a * 2 - 7 + num;
            if (nn < 0 || nn > MAX_NUM_MARKER - 1) {
                if (make_error) return false;
                nn = num;
            }
            for (int d = 5; d < DEPTH; ++d) at(x, y, d) = 0;
            at(x, y, 5 + nn) = 1;
        } else {
            return false;
        }
        hist.push_back(s);
        return true;
    }
};

// ---- interpreter ------------------------------------------------------------------------------
struct R { int n; bool s; bool c; };

static R run(const Node* nd, World& k, int n);

static R run_cond(const Node* nd, World& k, int n) {
    switch (nd->kind) {
        case K_COND_WRAP:                                       // r_cond1
            if (n > MAX_FUNC_CALL) return {n, false, false};
            return run_cond(nd->a, k, n);
        case K_COND_NOT: {                                      // r_cond2
            if (n > MAX_FUNC_CALL) return {n, false, false};
            R r = run_cond(nd->a, k, n);
            r.c = !r.c;
            return r;
        }
        case K_COND_PRIM: {
            if (n > MAX_FUNC_CALL) return {n, false, false};
            bool c = k.perceive(nd->arg);
            return {n, !k.broken, c};
        }
        default: return {n, false, false};
    }
}

static R run(const Node* nd, World& k, int n) {
    switch (nd->kind) {
        case K_PROG:                                            // r_prog
        case K_STMT:                                            // r_stmt
            if (n > MAX_FUNC_CALL) return {n, false, false};
            return run(nd->b, k, n + 1);
        case K_SEQ: {                                           // r_stmt_stmt
            if (n > MAX_FUNC_CALL) return {n, false, false};
            R r = run(nd->a, k, n + 1);
            if (!r.s) return r;
            if (r.n > MAX_FUNC_CALL) return {r.n, false, false};
            return run(nd->b, k, r.n);
        }
        case K_IF: {                                            // r_if
            if (n > MAX_FUNC_CALL) return {n, false, false};
            R r = run_cond(nd->a, k, n + 1);
            if (!r.s) return {r.n, false, false};
            if (r.c) return run(nd->b, k, r.n);
            return {r.n, true, false};
        }
        case K_IFELSE: {                                        // r_ifelse
            if (n > MAX_FUNC_CALL) return {n, false, false};
            R r = run_cond(nd->a, k, n + 1);
            if (!r.s) return {r.n, false, false};
            return run(r.c ? nd->b : nd->c, k, r.n);
        }
        case K_WHILE: {                                         // r_while
            if (n > MAX_FUNC_CALL) return {n, false, false};
            R r = run_cond(nd->a, k, n);
            if (!r.s) return {r.n, false, false};
            bool s = true;
            int nn = r.n;
            bool c = r.c;
            while (c) {
                R b = run(nd->b, k, nn);
                nn = b.n; s = b.s;
                if (!s) return {nn, false, false};
                R cc = run_cond(nd->a, k, nn);
                nn = cc.n; s = cc.s; c = cc.c;
                if (!s) return {nn, false, false};
            }
            return {nn, s, false};
        }
        case K_REPEAT: {                                        // r_repeat
            if (n > MAX_FUNC_CALL) return {n, false, false};
            int nn = n + 1;
            bool s = true;
            for (int i = 0; i < nd->arg; ++i) {
                R b = run(nd->b, k, nn);
                nn = b.n; s = b.s;
                if (!s) return {nn, false, false};
            }
            return {nn, s, false};
        }
        case K_ACTION:                                          // r_action1..5
            if (n > MAX_FUNC_CALL) return {n, false, false};
            return {n, k.act(nd->arg), false};
        default:                                                // cste / cond at statement position:
            return {n, false, false};                           // the reference would raise
    }
}

// ---- canonical form for the exact-program comparison (dsl_enum_program.py) --------------------------
// The reference flattens a program into a token list (WHILE = 100 copies of `if cond body`, REPEAT
// = n copies, IFELSE with equal branches collapses, `not not` cancels, noMarkersPresent = not
// markersPresent) and compares lists.  The lists grow as 100^depth, so they are compared through
// (length, two polynomial hashes) built compositionally instead of being materialised.
struct Flat {
    unsigned long long len; uint64_t h1, h2;
    bool operator==(const Flat& o) const { return len == o.len && h1 == o.h1 && h2 == o.h2; }
};
constexpr uint64_t M1 = 1000000007ULL, M2 = 998244353ULL, P1 = 911382323ULL, P2 = 972663749ULL;
static uint64_t powmod(uint64_t b, unsigned long long e, uint64_t m) {
    uint64_t r = 1; b %= m;
    while (e) { if (e & 1) r = r * b % m; b = b * b % m; e >>= 1; }
    return r;
}
static Flat fempty() { return {0, 0, 0}; }
static Flat fone(int word) { return {1, (uint64_t)(word + 1) % M1, (uint64_t)(word + 1) % M2}; }
static Flat fcat(const Flat& a, const Flat& b) {
    return {a.len + b.len, (a.h1 * powmod(P1, b.len, M1) + b.h1) % M1, (a.h2 * powmod(P2, b.len, M2) + b.h2) % M2};
}
static Flat frep(const Flat& a, int times) {
    Flat out = fempty(), base = a;
    while (times) { if (times & 1) out = fcat(out, base); base = fcat(base, base); times >>= 1; }
    return out;
}
enum Word { W_IF = 0, W_NOT, W_FRONT, W_LEFT, W_RIGHT, W_MARKERS, W_MOVE, W_TLEFT, W_TRIGHT, W_PICK, W_PUT };
struct CondFlat { bool neg; int prim; };   // [not] primitive
static CondFlat flat_cond(const Node* nd) {
    switch (nd->kind) {
        case K_COND_WRAP: return flat_cond(nd->a);
        case K_COND_NOT: { CondFlat c = flat_cond(nd->a); c.neg = !c.neg; return c; }
        default:
            if (nd->arg == 4) return {true, W_MARKERS};          // noMarkersPresent
            return {false, W_FRONT + nd->arg};
    }
}
static Flat cond_words(const CondFlat& c) {
    return c.neg ? fcat(fone(W_NOT), fone(c.prim)) : fone(c.prim);
}
static Flat flat(const Node* nd) {
    switch (nd->kind) {
        case K_PROG: case K_STMT: return flat(nd->b);
        case K_SEQ: return fcat(flat(nd->a), flat(nd->b));
        case K_IF: return fcat(fcat(fone(W_IF), cond_words(flat_cond(nd->a))), flat(nd->b));
        case K_IFELSE: {
            Flat s1 = flat(nd->b), s2 = flat(nd->c);
            if (s1 == s2) return s1;
            CondFlat c = flat_cond(nd->a), e = c;
            e.neg = !e.neg;
            Flat out = fcat(fcat(fone(W_IF), cond_words(c)), s1);
            return fcat(out, fcat(fcat(fone(W_IF), cond_words(e)), s2));
        }
        case K_WHILE:
            return frep(fcat(fcat(fone(W_IF), cond_words(flat_cond(nd->a))), flat(nd->b)), 100);
        case K_REPEAT: return frep(flat(nd->b), nd->arg);
        case K_ACTION: return fone(W_MOVE + nd->arg);
        default: return fempty();
    }
}

static bool statement_like(int sym) {
    return sym == S_PROG || sym == S_STMT || sym == S_STMT_STMT || sym == S_WHILE || sym == S_REPEAT ||
           sym == S_ACTION || sym == S_IF || sym == S_IFELSE;
}

}  // namespace karel
}  // namespace d2p

using namespace d2p;
using namespace d2p::karel;

extern "C" int d2p_karel_check_syntax(const int* tokens, int len) {
    if (!tokens && len > 0) return 0;
    Arena ar;
    return parse(tokens, len, ar, nullptr) != nullptr ? 1 : 0;
}

extern "C" int d2p_karel_execute(const int* tokens, int len, const unsigned char* state0, int h, int w,
                                 int make_error, int max_states, unsigned char* s_h, int* n_states) {
    D2P_REQUIRE(tokens && state0 && n_states && h > 0 && w > 0, "karel execute: bad arguments");
    Arena ar;
    int root = 0;
    Node* prog = parse(tokens, len, ar, &root);
    *n_states = 0;
    if (!prog) return -1;
    if (!statement_like(root)) return 0;
    World k;
    k.h = h; k.w = w; k.make_error = make_error != 0;
    const size_t fs = (size_t)h * w * DEPTH;
    k.s.resize(fs);
    for (size_t i = 0; i < fs; ++i) k.s[i] = state0[i] ? 1 : 0;   // astype(bool)
    k.hist.push_back(k.s);
    R r = run(prog, k, 0);
    if (!r.s || k.broken) return 0;
    *n_states = (int)k.hist.size();
    if (s_h)
        for (int t = 0; t < (int)k.hist.size() && t < max_states; ++t)
            std::memcpy(s_h + (size_t)t * fs, k.hist[t].data(), fs);
    return 1;
}

extern "C" int d2p_karel_programs_equal(const int* a, int la, const int* b, int lb) {
    Arena ar;
    int ra = 0, rb = 0;
    Node* pa = parse(a, la, ar, &ra);
    Node* pb = parse(b, lb, ar, &rb);
    if (!pa || !pb || ra != S_PROG || rb != S_PROG) return -1;
    return flat(pa) == flat(pb) ? 1 : 0;
}

// Batch metrics of one evaluation step.  tokens [B, L] (argmax of the decoder), lens [B],
// is_same_seq [B] (prediction == ground truth incl. length); demos [B, k, T, h, w, 16] u8,
// demo_len [B, k].  Outputs: is_correct_syntax [B], is_correct_execution [B, k],
// num_correct_execution [B].
extern "C" int d2p_karel_eval_batch(const int* tokens, const int* lens, const unsigned char* is_same_seq, int B,
                                    int L, const unsigned char* demos, const int* demo_len, int k, int T, int h,
                                    int w, int make_error, float* is_correct_syntax,
                                    float* is_correct_execution, float* num_correct_execution, int nthreads) {
    D2P_REQUIRE(tokens && lens && is_same_seq && demos && demo_len && is_correct_syntax && is_correct_execution &&
                num_correct_execution, "karel eval: null buffer");
    D2P_REQUIRE(B > 0 && L > 0 && k > 0 && T > 0 && h > 0 && w > 0, "karel eval: bad dims");
    const size_t fs = (size_t)h * w * DEPTH;
    auto work = [&](int b0, int b1) {
        std::vector<unsigned char> exe((size_t)T * fs);
        for (int b = b0; b < b1; ++b) {
            int len = lens[b] < 0 ? 0 : (lens[b] > L ? L : lens[b]);
            const int* tk = tokens + (size_t)b * L;
            const bool same = is_same_seq[b] != 0;
            const bool syntax = same || d2p_karel_check_syntax(tk, len) == 1;
            is_correct_syntax[b] = syntax ? 1.f : 0.f;
            float nc = 0.f;
            for (int i = 0; i < k; ++i) {
                const unsigned char* demo = demos + ((size_t)b * k + i) * T * fs;
                bool ok = same;
                if (!same) {
                    int n_states = 0;
                    std::fill(exe.begin(), exe.end(), 0);
                    // only run when the program is new and parses (model_full.py:761)
                    if (syntax && d2p_karel_execute(tk, len, demo, h, w, make_error, T, exe.data(), &n_states) != 1) {
                        n_states = 0;
                        std::fill(exe.begin(), exe.end(), 0);
                    }
                    bool eq = n_states == demo_len[(size_t)b * k + i];
                    for (size_t j = 0; eq && j < (size_t)T * fs; ++j) eq = (demo[j] != 0) == (exe[j] != 0);
                    ok = eq;
                }
                is_correct_execution[(size_t)b * k + i] = ok ? 1.f : 0.f;
                nc += ok ? 1.f : 0.f;
            }
            num_correct_execution[b] = nc;
        }
    };
    int nt = nthreads > 0 ? nthreads : (int)std::thread::hardware_concurrency();
    if (nt < 1) nt = 1;
    if (nt > B) nt = B;
    if (nt == 1) { work(0, B); return 0; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; ++t) th.emplace_back(work, (int)((long long)B * t / nt), (int)((long long)B * (t + 1) / nt));
    for (auto& x : th) x.join();
    return 0;
}
