// ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by, or executed
// from the product path (grample_b200/, include/).  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
//
// CPU float64 restatement of the reference's model layer (Go package `model`).
// The reference is pure Go and no Go toolchain exists in this image, so the reference
// itself cannot be compiled (oracle/_ref is therefore not applicable); this file
// restates the algorithm line range by line range.  Citations are relative to
// /root/reference.
//
// Pinned by the reference's own known-answer tests (transcribed in tests/test_oracle_*.py):
//   function_test.go:81-157, 192-251   table index order, log identity, AddValue
//   variable_iter_test.go:40-138       enumeration order, honorFixed
//   error_test.go:11-119               Hellinger / JSD / abs-error suite
//   uai_test.go:31-208                 UAI / evidence / MAR parsers
#pragma once
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace oracle {

// Go's (value, error) convention is restated with exceptions; every throw site
// corresponds to a `return ..., errors.Errorf(...)` in the reference.
struct Error : std::runtime_error {
    using std::runtime_error::runtime_error;
};

// ---------------------------------------------------------------- Variable
// model/variable.go:10-18
struct Variable {
    int id = 0;
    std::string name;
    int card = 0;
    int fixed_val = -1;  // -1 = no evidence
    std::vector<double> marginal;
    std::map<std::string, double> state;
    bool collapsed = false;

    // model/variable.go:106-147
    void norm_marginal() {
        if (card != (int)marginal.size()) throw Error("can not norm: Card != len(Marginal)");
        if (card < 1) return;
        if (card == 1) marginal[0] = 1.0;
        double sum = 0.0;
        for (double p : marginal) sum += p;
        const double EPS = 1e-8;
        if (std::fabs(sum - 1.0) < EPS) return;
        if (std::fabs(sum) < EPS) {
            double p = 1.0 / (double)card;
            for (auto& m : marginal) m = p;
            return;
        }
        for (auto& m : marginal) m = m / sum;
    }

    // model/variable.go:76-103
    void check() const {
        if (card != (int)marginal.size()) throw Error("Variable Card != len(M)");
        if (fixed_val != -1 && (fixed_val < 0 || fixed_val >= card))
            throw Error("Variable has fixed val that does not match card");
        if (card > 0) {
            double sum = 0.0;
            for (double p : marginal) sum += p;
            if (std::fabs(sum - 1.0) >= 1e-8) throw Error("Variable has marginal dist with sum != 1");
        }
    }
};

// model/variable.go:167-189 (Excel-column style names; only used for messages)
inline std::string letter26(int n) {
    if (n == 0) return "A";
    n++;
    std::string digits;
    while (n > 0) {
        int rem = (n - 1) % 26;
        n = (n - 1) / 26;
        digits.insert(digits.begin(), (char)('A' + rem));
    }
    return digits;
}

// model/variable.go:22-51
inline Variable new_variable(int index, int card) {
    if (index < 0) throw Error("Invalid index for variable");
    if (card < 1) throw Error("Invalid card for variable");
    Variable v;
    v.id = index;
    v.card = card;
    v.fixed_val = -1;
    v.marginal.assign(card, 0.0);
    v.name = letter26(index);
    v.norm_marginal();  // all-zero -> uniform 1/card
    return v;
}

// ---------------------------------------------------------------- Function
// model/function.go:37-42.  The reference stores *Variable pointers; only ID and Card
// are ever read through them on this path, so ids + cards are kept instead.
constexpr int kMaxTabSize = 1 << 23;  // model/function.go:59

struct Function {
    std::string name;
    std::vector<int> vars;   // variable ids, most significant first
    std::vector<int> cards;  // cardinality of each scope variable
    std::vector<double> table;
    bool is_log = false;

    // model/function.go:180-202 — last scope variable is the fastest digit
    long calc_index(const int* values, size_t n) const {
        if (n != vars.size()) throw Error("Value vector does not match variables");
        long digit = 1, location = 0;
        for (long i = (long)n - 1; i >= 0; i--) {
            int val = values[i];
            int card = cards[i];
            if (val < 0 || val >= card) throw Error("Value invalid for cardinality");
            location += digit * val;
            digit *= card;
        }
        return location;
    }

    // model/function.go:146-157
    double eval(const int* values, size_t n) const {
        long i = calc_index(values, n);
        if (i < 0 || i >= (long)table.size()) throw Error("Could not find table entry");
        return table[i];
    }
    double eval(const std::vector<int>& v) const { return eval(v.data(), v.size()); }

    // model/function.go:161-177
    void add_value(const int* values, size_t n, double inc) {
        if (is_log) throw Error("Can not AddValue if function is already in log space");
        long i = calc_index(values, n);
        if (i < 0 || i >= (long)table.size()) throw Error("Could not find table entry");
        table[i] += inc;
    }

    // model/function.go:126-142 — NOTE: v += eps (not max), only when v < eps
    void use_log_space() {
        if (is_log) throw Error("IsLog already set - double-call detected");
        const double eps = 1e-6;
        for (auto& t : table) {
            double v = t;
            if (v < eps) v += eps;
            t = std::log(v);
        }
        is_log = true;
    }

    // model/function.go:110-122
    void check() const {
        long ts = 0;
        for (size_t i = 0; i < cards.size(); i++) ts = (i == 0 ? 1 : ts) * cards[i];
        if (ts < 1) throw Error("Function is invalid - can not calculate table size");
        if (ts != (long)table.size()) throw Error("Function table size mismatch");
    }
};

// model/function.go:47-57, 62-89
inline Function new_function(int index, const std::vector<int>& var_ids, const std::vector<int>& cards) {
    if (index < 0) throw Error("Invalid index for function");
    Function f;
    f.name = "func-" + std::to_string(index);
    if (var_ids.empty()) throw Error("Empty variable list for function is invalid");
    long ts = 0;
    for (size_t i = 0; i < cards.size(); i++) {
        if (i == 0) ts = 1;
        ts *= cards[i];
        if (ts > (long)kMaxTabSize * 64) break;  // avoid overflow; still > cap
    }
    if (ts < 1) throw Error("Function is invalid - could not calculate table size");
    if (ts > kMaxTabSize) throw Error("Function table size exceeds 1<<23");
    f.vars = var_ids;
    f.cards = cards;
    f.table.assign(ts, 0.0);
    return f;
}

// ---------------------------------------------------------------- Model
// model/model.go:24-49, 115-157
struct Model {
    std::string type;
    std::string name;
    std::vector<Variable> vars;
    std::vector<Function> funcs;

    Model clone() const { return *this; }  // deep copy (value semantics)

    void check() const {
        if (type != "BAYES" && type != "MARKOV") throw Error("Unknown model type " + type);
        std::map<int, bool> seen;
        int fix = 0;
        for (auto& v : vars) {
            v.check();
            if (seen.count(v.id)) throw Error("Duplicate Id for Var");
            seen[v.id] = true;
            if (v.fixed_val > -1) fix++;
        }
        if (fix >= (int)vars.size()) throw Error("all vars are fixed!");
        std::map<std::string, bool> names;
        for (auto& f : funcs) {
            f.check();
            names[f.name] = true;
        }
        if (names.size() != funcs.size()) throw Error("function names are not unique");
    }
};

// ---------------------------------------------------------------- VariableIter
// model/variable_iter.go:15-74 — odometer, last variable fastest, honours FixedVal
struct VariableIter {
    std::vector<int> cards, fixed, last;
    bool honor_fixed;

    VariableIter(const std::vector<const Variable*>& src, bool honor) : honor_fixed(honor) {
        if (src.empty()) throw Error("At least one variable required for iteration");
        for (auto* v : src) {
            cards.push_back(v->card);
            fixed.push_back(v->fixed_val);
            last.push_back((honor && v->fixed_val >= 0) ? v->fixed_val : 0);
        }
    }
    void val(std::vector<int>& cur) const {
        if (cur.size() < last.size()) throw Error("Dest buffer too small");
        for (size_t i = 0; i < last.size(); i++) cur[i] = last[i];
    }
    bool next() {
        for (long i = (long)cards.size() - 1; i >= 0; i--) {
            if (honor_fixed && fixed[i] >= 0) {
                last[i] = fixed[i];
                continue;
            }
            int prop = last[i] + 1;
            if (prop < cards[i]) {
                last[i] = prop;
                return true;
            }
            last[i] = 0;
        }
        return false;
    }
};

// ---------------------------------------------------------------- distance measures
// model/error.go:81-249.  All return 0 if either side is fixed.
enum Measure { kMaxAbs = 0, kMeanAbs = 1, kHellinger = 2, kJS = 3 };

inline void totals(const Variable& a, const Variable& b, double& t1, double& t2) {
    t1 = 0.0; t2 = 0.0;
    for (int c = 0; c < a.card; c++) { t1 += a.marginal[c]; t2 += b.marginal[c]; }
    const double eps = 1e-12;
    if (t1 < eps) t1 = eps;
    if (t2 < eps) t2 = eps;
}
inline double max_abs_diff(const Variable& a, const Variable& b) {  // error.go:81-114
    if (a.fixed_val >= 0 || b.fixed_val >= 0) return 0.0;
    double t1, t2; totals(a, b, t1, t2);
    double mx = 0.0;
    for (int c = 0; c < a.card; c++) {
        double e = std::fabs(a.marginal[c] / t1 - b.marginal[c] / t2);
        if (c == 0 || e > mx) mx = e;
    }
    return mx;
}
inline double mean_abs_diff(const Variable& a, const Variable& b) {  // error.go:117-151
    if (a.fixed_val >= 0 || b.fixed_val >= 0) return 0.0;
    if (a.card < 1) return 0.0;
    double t1, t2; totals(a, b, t1, t2);
    double s = 0.0;
    for (int c = 0; c < a.card; c++) s += std::fabs(a.marginal[c] / t1 - b.marginal[c] / t2);
    return s / (double)a.card;
}
inline double hellinger_diff(const Variable& a, const Variable& b) {  // error.go:158-190
    if (a.fixed_val >= 0 || b.fixed_val >= 0) return 0.0;
    double t1, t2; totals(a, b, t1, t2);
    double s = 0.0;
    for (int c = 0; c < a.card; c++) {
        double d = std::sqrt(a.marginal[c] / t1) - std::sqrt(b.marginal[c] / t2);
        s += d * d;
    }
    return std::sqrt(s) / std::sqrt(2.0);
}
inline double kl_div(const std::vector<double>& p, const std::vector<double>& q) {  // error.go:197-212
    const double eps = 1e-12;
    double d = 0.0;
    for (size_t i = 0; i < p.size(); i++) {
        double p1 = p[i] < eps ? eps : p[i];
        double p2 = q[i] < eps ? eps : q[i];
        d += p1 * std::log2(p1 / p2);
    }
    return d;
}
inline double js_divergence(const Variable& a, const Variable& b) {  // error.go:216-249
    if (a.fixed_val >= 0 || b.fixed_val >= 0) return 0.0;
    double t1, t2; totals(a, b, t1, t2);
    std::vector<double> p1(a.card), p2(a.card), mid(a.card);
    for (int i = 0; i < a.card; i++) {
        p1[i] = a.marginal[i] / t1;
        p2[i] = b.marginal[i] / t2;
        mid[i] = (p1[i] + p2[i]) * 0.5;
    }
    return 0.5 * (kl_div(p1, mid) + kl_div(p2, mid));
}
inline double measure(int m, const Variable& a, const Variable& b) {
    switch (m) {
        case kMaxAbs: return max_abs_diff(a, b);
        case kMeanAbs: return mean_abs_diff(a, b);
        case kHellinger: return hellinger_diff(a, b);
        case kJS: return js_divergence(a, b);
    }
    throw Error("unknown measure");
}

// model/error.go:15-78
struct ErrorSuite {
    double mean_mean_abs = 0, mean_max_abs = 0, mean_hellinger = 0, mean_js = 0;
    double max_mean_abs = 0, max_max_abs = 0, max_hellinger = 0, max_js = 0;
};
inline ErrorSuite new_error_suite(const std::vector<Variable>& v1, const std::vector<Variable>& v2) {
    if (v1.size() != v2.size()) throw Error("Variable count mismatch");
    int cnt = 0;
    for (size_t i = 0; i < v1.size(); i++) {
        if (v1[i].card != v2[i].card) throw Error("Variable card mismatch");
        if (v1[i].fixed_val < 0 && v2[i].fixed_val < 0) cnt++;
    }
    if (cnt < 1) throw Error("No un-fixed vars to score");
    ErrorSuite es;
    for (size_t i = 0; i < v1.size(); i++) {
        double d;
        d = mean_abs_diff(v1[i], v2[i]); es.mean_mean_abs += d; es.max_mean_abs = std::fmax(d, es.max_mean_abs);
        d = max_abs_diff(v1[i], v2[i]);  es.mean_max_abs += d;  es.max_max_abs = std::fmax(d, es.max_max_abs);
        d = hellinger_diff(v1[i], v2[i]); es.mean_hellinger += d; es.max_hellinger = std::fmax(d, es.max_hellinger);
        d = js_divergence(v1[i], v2[i]); es.mean_js += d; es.max_js = std::fmax(d, es.max_js);
    }
    double fc = (double)cnt;
    es.mean_mean_abs /= fc; es.mean_max_abs /= fc; es.mean_hellinger /= fc; es.mean_js /= fc;
    return es;
}

// ---------------------------------------------------------------- UAI reader
// model/uai.go:20-50 — drop blank/comment ('c') lines; optionally skip to reqPrefix
inline std::string trim(const std::string& s) {
    size_t a = 0, b = s.size();
    auto ws = [](unsigned char c) { return c == ' ' || c == '\t' || c == '\r' || c == '\n' || c == '\v' || c == '\f'; };
    while (a < b && ws(s[a])) a++;
    while (b > a && ws(s[b - 1])) b--;
    return s.substr(a, b - a);
}
inline std::string uai_preprocess(const std::string& data, const std::string& req_prefix, int& line_count) {
    std::vector<std::string> kept;
    bool start_found = req_prefix.empty();
    size_t pos = 0;
    while (pos <= data.size()) {
        size_t nl = data.find('\n', pos);
        std::string ln = trim(data.substr(pos, nl == std::string::npos ? std::string::npos : nl - pos));
        pos = (nl == std::string::npos) ? data.size() + 1 : nl + 1;
        if (ln.empty() || ln[0] == 'c') continue;
        if (!start_found) {
            if (ln.compare(0, req_prefix.size(), req_prefix) == 0) start_found = true;
            else continue;
        }
        kept.push_back(ln);
    }
    line_count = (int)kept.size();
    std::string out;
    for (size_t i = 0; i < kept.size(); i++) {
        if (i) out += "\n";
        out += kept[i];
    }
    return out;
}

// model/reader.go:10-49
struct FieldReader {
    size_t pos = 0;
    std::vector<std::string> fields;
    explicit FieldReader(const std::string& text) {
        std::istringstream ss(text);
        std::string tok;
        while (ss >> tok) fields.push_back(tok);
    }
    std::string read() {
        if (pos >= fields.size()) throw Error("EOF");
        return fields[pos++];
    }
    int read_int() {
        std::string s = read();
        char* end = nullptr;
        long v = std::strtol(s.c_str(), &end, 10);
        if (end == s.c_str() || *end != '\0') throw Error("invalid int: " + s);
        return (int)v;
    }
    double read_float() {
        std::string s = read();
        char* end = nullptr;
        double v = std::strtod(s.c_str(), &end);
        if (end == s.c_str() || *end != '\0') throw Error("invalid float: " + s);
        return v;
    }
};

// model/uai.go:53-179
inline Model read_model(const std::string& data) {
    if (data.size() < 15) throw Error("Invalid data buffer: len<15");
    int lc = 0;
    std::string text = uai_preprocess(data, "", lc);
    if (lc < 1) throw Error("No lines found in file");
    FieldReader fr(text);
    if (fr.fields.size() < 6) throw Error("Invalid data: <6 fields");
    Model m;
    m.type = fr.read();
    if (m.type != "BAYES" && m.type != "MARKOV") throw Error("Unknown model type " + m.type);
    int nv = fr.read_int();
    if (nv < 1) throw Error("Invalid variable count");
    for (int i = 0; i < nv; i++) {
        int card = fr.read_int();
        if (card < 1) throw Error("Invalid card");
        m.vars.push_back(new_variable(i, card));
    }
    int nf = fr.read_int();
    if (nf < 1) throw Error("Invalid Clique count");
    for (int i = 0; i < nf; i++) {
        int k = fr.read_int();
        if (k < 1) throw Error("Invalid variable count (<1) for Clique");
        std::vector<int> ids(k), cards(k);
        for (int j = 0; j < k; j++) {
            int vi = fr.read_int();
            if (vi < 0 || vi >= nv) throw Error("Invalid var idx for Clique");
            ids[j] = vi;
            cards[j] = m.vars[vi].card;
        }
        m.funcs.push_back(new_function(i, ids, cards));
    }
    for (auto& f : m.funcs) {
        int ts = fr.read_int();
        if (ts != (int)f.table.size()) throw Error("Read table size != Clique size on " + f.name);
        for (int t = 0; t < ts; t++) f.table[t] = fr.read_float();
    }
    return m;
}

// model/uai.go:183-249
inline void apply_evidence(const std::string& data, Model& m) {
    int lc = 0;
    std::string text = uai_preprocess(data, "", lc);
    if (lc < 1) throw Error("Invalid data buffer: there is no data");
    if (lc > 2) throw Error("only understand evidence files with 1 or 2 lines");
    FieldReader fr(text);
    if (fr.fields.empty()) throw Error("Invalid data: found no fields");
    if (lc == 2) {
        int sc = fr.read_int();
        if (sc == 0) return;
        if (sc > 1) throw Error("only single sample evidence currently supported");
    }
    int n = fr.read_int();
    if (n < 1) return;
    for (int i = 0; i < n; i++) {
        int idx = fr.read_int();
        if (idx < 0 || idx >= (int)m.vars.size()) throw Error("Read incorrect variable index");
        Variable& v = m.vars[idx];
        if (v.fixed_val != -1) throw Error("variable had previous fixedval");
        int val = fr.read_int();
        if (val < 0 || val >= v.card) throw Error("Read invalid value for variable");
        v.fixed_val = val;
    }
}

// model/uai.go:252-332 ; model/solution.go:16-65
struct Solution {
    std::vector<Variable> vars;
    void check(const Model& m) const {
        for (auto& v : vars) v.check();
        if (vars.size() != m.vars.size()) throw Error("Solution var count != model var count");
    }
    ErrorSuite error(const std::vector<Variable>& other) const { return new_error_suite(vars, other); }
};
inline Solution read_marg_solution(const std::string& data) {
    if (data.size() < 11) throw Error("Invalid data buffer: len<11");
    int lc = 0;
    std::string text = uai_preprocess(data, "MAR", lc);
    if (lc < 1) throw Error("No lines in file");
    FieldReader fr(text);
    if (fr.fields.size() < 4) throw Error("Invalid data: <4 fields");
    std::string t = fr.read();
    if (t != "MAR") throw Error("Unknown solution file type " + t);
    int nv = fr.read_int();
    if (nv < 1) throw Error("Invalid variable count");
    Solution sol;
    for (int i = 0; i < nv; i++) {
        int card = fr.read_int();
        if (card < 1) throw Error("Invalid card");
        Variable v = new_variable(i, card);
        for (int k = 0; k < card; k++) {
            double p = fr.read_float();
            if (p < 0.0 || p > 1.0) throw Error("Invalid marg prob");
            v.marginal[k] = p;
        }
        v.norm_marginal();
        sol.vars.push_back(v);
    }
    return sol;
}

inline std::string slurp(const std::string& path) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw Error("Could not READ " + path);
    std::ostringstream ss;
    ss << f.rdbuf();
    return ss.str();
}

// model/model.go:52-112
inline Model model_from_buffer(const std::string& data) {
    Model m = read_model(data);
    m.check();
    return m;
}
inline Model model_from_file(const std::string& filename, bool use_evidence) {
    Model m = model_from_buffer(slurp(filename));
    size_t dot = filename.find_last_of('.');
    m.name = (dot == std::string::npos) ? filename : filename.substr(0, dot);
    if (use_evidence) {
        for (auto& v : m.vars) v.fixed_val = -1;
        apply_evidence(slurp(filename + ".evid"), m);
    }
    return m;
}

}  // namespace oracle
