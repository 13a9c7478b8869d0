// Array-native newick ingest: every ';'-terminated tree of a text buffer -> ONE forest in level order (the node
// numbering of the engine and of FlatForest), without building Python node objects.
//
// Replaces pastml.tree.read_forest / read_tree + the ete3 traversal for large trees (pastml/tree.py:176-222):
// labels of internal nodes are node names (ete3 format 1), a missing branch length is ete3's default (1.0), the
// root's is 0 unless given, negative branch lengths are set to zero (tree.py:194-198), quoted labels keep their
// content, [comments] (NHX fields included) are skipped.  pml_newick_name_nodes applies the naming rule of
// pastml.tree.name_tree (tree.py:76-105) to one tree of the forest.
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <string_view>
#include <unordered_map>
#include <unordered_set>
#include <vector>

#include "../../include/pastml_b200.h"

namespace {

struct RawNode {
    int32_t parent;       // index in parse order, -1 for roots
    double dist;
    std::string name;
    std::vector<int32_t> children;
};

}  // namespace

struct pml_newick {
    // level order
    std::vector<int32_t> parent;
    std::vector<double> dist;
    std::vector<std::string> names;
    int64_t n_trees = 0;
    std::string error;
};

namespace {

thread_local std::string g_newick_error;

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r' || c == '\f' || c == '\v'; }

std::string strip(const std::string &s) {
    size_t b = 0, e = s.size();
    while (b < e && is_space(s[b])) ++b;
    while (e > b && is_space(s[e - 1])) --e;
    return s.substr(b, e - b);
}

// label := [name][:dist][comment]*  -- returns the position after the label, or -1 on error
int64_t read_label(const char *t, int64_t pos, int64_t n, RawNode &target, std::string &err) {
    std::string name, num;
    bool seen_colon = false;
    while (pos < n) {
        const char c = t[pos];
        if (c == '(' || c == ')' || c == ',' || c == ';') break;
        if (c == '[') {
            const void *end = memchr(t + pos, ']', (size_t) (n - pos));
            if (!end) { err = "Unclosed comment in newick"; return -1; }
            pos = (const char *) end - t + 1;
            continue;
        }
        if ((c == '"' || c == '\'') && !seen_colon && name.empty()) {
            const void *end = memchr(t + pos + 1, c, (size_t) (n - pos - 1));
            if (!end) { err = "Unclosed quote in newick"; return -1; }
            const int64_t e = (const char *) end - t;
            name.append(t + pos + 1, (size_t) (e - pos - 1));
            pos = e + 1;
            continue;
        }
        if (c == ':' && !seen_colon) {
            seen_colon = true;
            ++pos;
            continue;
        }
        (seen_colon ? num : name).push_back(c);
        ++pos;
    }
    name = strip(name);
    if (!name.empty()) target.name = name;
    if (seen_colon) {
        num = strip(num);
        if (!num.empty()) {
            char *endp = nullptr;
            const double v = strtod(num.c_str(), &endp);
            if (endp == num.c_str() || *endp != '\0') { err = "Bad branch length \"" + num + "\""; return -1; }
            target.dist = v;
        }
    }
    return pos;
}

// parses one tree (text without the trailing ';') into `nodes` (appended); returns the root index or -1
int32_t parse_one(const char *t, int64_t n, std::vector<RawNode> &nodes, std::string &err) {
    while (n > 0 && is_space(t[0])) { ++t; --n; }
    while (n > 0 && is_space(t[n - 1])) --n;
    if (n == 0) { err = "Empty newick string"; return -1; }
    const int32_t root = (int32_t) nodes.size();
    nodes.push_back(RawNode{-1, 0., std::string(), {}});
    if (t[0] != '(') {   // a single-node tree: "A;" or "A:1;"
        const int64_t pos = read_label(t, 0, n, nodes[root], err);
        if (pos < 0) return -1;
        if (pos != n) { err = "Unexpected newick format"; return -1; }
        return root;
    }
    int32_t node = root;
    int64_t depth = 0, pos = 0;
    bool expecting_label = false;
    while (pos < n) {
        const char c = t[pos];
        if (c == '(') {
            if (expecting_label) { err = "Unexpected \"(\" in newick"; return -1; }
            const int32_t child = (int32_t) nodes.size();
            nodes.push_back(RawNode{node, 1., std::string(), {}});
            nodes[node].children.push_back(child);
            node = child;
            ++depth;
            ++pos;
            if (pos < n && t[pos] != '(' && t[pos] != ')' && t[pos] != ',') {
                pos = read_label(t, pos, n, nodes[node], err);
                if (pos < 0) return -1;
            }
        } else if (c == ',') {
            if (depth == 0) { err = "Comma outside parentheses in newick"; return -1; }
            const int32_t par = nodes[node].parent;
            const int32_t child = (int32_t) nodes.size();
            nodes.push_back(RawNode{par, 1., std::string(), {}});
            nodes[par].children.push_back(child);
            node = child;
            expecting_label = false;
            ++pos;
            if (pos < n && t[pos] != '(' && t[pos] != ')' && t[pos] != ',') {
                pos = read_label(t, pos, n, nodes[node], err);
                if (pos < 0) return -1;
            }
        } else if (c == ')') {
            if (depth == 0) { err = "Unbalanced parentheses in newick"; return -1; }
            node = nodes[node].parent;
            --depth;
            ++pos;
            if (node == root) nodes[node].dist = 0.;
            pos = read_label(t, pos, n, nodes[node], err);
            if (pos < 0) return -1;
            expecting_label = false;
        } else if (is_space(c)) {
            ++pos;
        } else {
            err = std::string("Unexpected character \"") + c + "\" in newick";
            return -1;
        }
    }
    if (depth != 0 || node != root) { err = "Unbalanced parentheses in newick"; return -1; }
    return root;
}

}  // namespace

extern "C" {

const char *pml_newick_last_error(void) { return g_newick_error.c_str(); }

int pml_newick_parse(const char *text, int64_t len, pml_newick **out) {
    if (!text || !out || len < 0) { g_newick_error = "null argument"; return PML_ERR_INVALID; }
    std::vector<RawNode> nodes;
    std::vector<int32_t> roots;
    // trees are separated by ';' (newlines inside a tree are dropped by the whitespace rule, tree.py:189)
    int64_t start = 0;
    std::string err;
    for (int64_t i = 0; i <= len; ++i) {
        if (i == len || text[i] == ';') {
            bool blank = true;
            for (int64_t k = start; k < i && blank; ++k) blank = is_space(text[k]);
            if (!blank) {
                if (i == len) { start = i + 1; continue; }   // text after the last ';' is ignored, as in read_forest
                const int32_t r = parse_one(text + start, i - start, nodes, err);
                if (r < 0) { g_newick_error = err; return PML_ERR_INVALID; }
                roots.push_back(r);
            }
            start = i + 1;
        }
    }
    if (roots.empty()) {
        // a single tree without the terminating ';'
        bool blank = true;
        for (int64_t k = 0; k < len && blank; ++k) blank = is_space(text[k]);
        if (blank) { g_newick_error = "Could not find any trees (in newick format)"; return PML_ERR_INVALID; }
        const int32_t r = parse_one(text, len, nodes, err);
        if (r < 0) { g_newick_error = err; return PML_ERR_INVALID; }
        roots.push_back(r);
    }
    // level order over the whole forest: all roots, then all depth-1 nodes (tree by tree, children left to right), ...
    pml_newick *nw = new pml_newick();
    nw->n_trees = (int64_t) roots.size();
    const size_t N = nodes.size();
    nw->parent.reserve(N);
    nw->dist.reserve(N);
    nw->names.reserve(N);
    std::vector<int32_t> frontier = roots, frontier_parent(roots.size(), -1), next, next_parent;
    while (!frontier.empty()) {
        next.clear();
        next_parent.clear();
        for (size_t k = 0; k < frontier.size(); ++k) {
            RawNode &rn = nodes[frontier[k]];
            const int32_t idx = (int32_t) nw->parent.size();
            nw->parent.push_back(frontier_parent[k]);
            nw->dist.push_back(rn.dist < 0. ? 0. : rn.dist);
            nw->names.push_back(std::move(rn.name));
            for (int32_t ch : rn.children) {
                next.push_back(ch);
                next_parent.push_back(idx);
            }
        }
        frontier.swap(next);
        frontier_parent.swap(next_parent);
    }
    *out = nw;
    return PML_OK;
}

int64_t pml_newick_num_nodes(const pml_newick *nw) { return nw ? (int64_t) nw->parent.size() : 0; }
int64_t pml_newick_num_trees(const pml_newick *nw) { return nw ? nw->n_trees : 0; }

int64_t pml_newick_names_bytes(const pml_newick *nw) {
    int64_t total = 0;
    if (nw)
        for (const std::string &s : nw->names) total += (int64_t) s.size();
    return total;
}

int pml_newick_fill(const pml_newick *nw, int32_t *parent, double *dist, int64_t *name_offsets, char *names) {
    if (!nw || !parent || !dist) { g_newick_error = "null argument"; return PML_ERR_INVALID; }
    const size_t N = nw->parent.size();
    memcpy(parent, nw->parent.data(), N * sizeof(int32_t));
    memcpy(dist, nw->dist.data(), N * sizeof(double));
    if (name_offsets && names) {
        int64_t off = 0;
        for (size_t i = 0; i < N; ++i) {
            name_offsets[i] = off;
            memcpy(names + off, nw->names[i].data(), nw->names[i].size());
            off += (int64_t) nw->names[i].size();
        }
        name_offsets[N] = off;
    }
    return PML_OK;
}

// name_tree (tree.py:76-105) on every tree of the forest: unnamed nodes and names used 10 times or more are
// replaced by t<i> / n<i> (preorder counter), the root becomes "root", collisions get the counter appended.
// Trees whose names are already unique and complete are left alone.  suffix: "" for a single tree, "_<i>" per tree
// of a forest is the caller's choice (acr.py:163-165 passes the tree index for forests).
int pml_newick_name_nodes(pml_newick *nw, int use_tree_suffix) {
    if (!nw) { g_newick_error = "null argument"; return PML_ERR_INVALID; }
    const size_t N = nw->parent.size();
    // children lists and tree ids from the level-ordered parent array
    std::vector<std::vector<int32_t>> children(N);
    std::vector<int32_t> tree_of(N, 0), roots;
    for (size_t i = 0; i < N; ++i) {
        const int32_t p = nw->parent[i];
        if (p < 0) {
            tree_of[i] = (int32_t) roots.size();
            roots.push_back((int32_t) i);
        } else {
            tree_of[i] = tree_of[p];
            children[p].push_back((int32_t) i);
        }
    }
    for (size_t t = 0; t < roots.size(); ++t) {
        const std::string suffix = (use_tree_suffix && roots.size() > 1) ? "_" + std::to_string(t) : std::string();
        // preorder of this tree
        std::vector<int32_t> order, stack{roots[t]};
        while (!stack.empty()) {
            const int32_t v = stack.back();
            stack.pop_back();
            order.push_back(v);
            for (size_t k = children[v].size(); k-- > 0;) stack.push_back(children[v][k]);
        }
        // names of this tree and how often each occurs (views into the stored names: nothing is renamed before the
        // counting is over, and the renaming pass below looks a name up before it overwrites it)
        std::unordered_map<std::string_view, int> seen;
        seen.reserve(order.size() * 2);
        for (int32_t v : order)
            if (!nw->names[v].empty()) seen[std::string_view(nw->names[v])]++;
        if (seen.size() == order.size()) continue;   // every node has its own name
        long long counter = 0;
        std::vector<std::string> renamed(order.size());
        std::unordered_set<std::string_view> taken;
        taken.reserve(order.size() * 2);
        std::string candidate, prefix;
        for (size_t k = 0; k < order.size(); ++k) {
            const int32_t v = order[k];
            const std::string &nm = nw->names[v];
            if (!nm.empty() && seen[std::string_view(nm)] < 10) prefix = nm;
            else prefix = (children[v].empty() ? "t" : "n") + std::to_string(counter) + suffix;
            candidate = (nw->parent[v] < 0) ? "root" + suffix : prefix;
            while (taken.count(std::string_view(candidate))) {
                candidate = prefix + std::to_string(counter) + suffix;
                ++counter;
            }
            renamed[k] = candidate;
            taken.insert(std::string_view(renamed[k]));   // renamed[] is never resized: the views stay valid
        }
        for (size_t k = 0; k < order.size(); ++k) nw->names[order[k]] = std::move(renamed[k]);
    }
    return PML_OK;
}

void pml_newick_free(pml_newick *nw) { delete nw; }


// ---- result tables
//
// The marginal-probability tables of a reconstruction (acr.py:45-73: DataFrame.to_csv(sep='\t', index_label='node')) for
// array-native results: "label \t v1 \t v2 ... \n" per node, every number in Python's repr() form -- the shortest
// digits that round-trip, fixed notation while the decimal exponent lies in (-4, 16], exponent notation with at least
// two exponent digits otherwise, an empty cell for NaN.  Host only; 2M numbers of a 100k-tip tree take 0.1 s instead of
// the seconds pandas spends converting floats to strings.
static char *append_repr(char *out, double x) {
    if (x != x) return out;                      // NaN: empty cell (pandas' na_rep)
    if (x == 0.) {
        if (std::signbit(x)) *out++ = '-';
        *out++ = '0'; *out++ = '.'; *out++ = '0';
        return out;
    }
    if (std::isinf(x)) {
        if (x < 0) *out++ = '-';
        *out++ = 'i'; *out++ = 'n'; *out++ = 'f';
        return out;
    }
    char sci[40];
    const auto res = std::to_chars(sci, sci + sizeof(sci) - 1, x, std::chars_format::scientific);   // d[.ddd]e[+-]XX, shortest
    *res.ptr = 0;
    const char *p = sci, *end = res.ptr;
    if (*p == '-') { *out++ = '-'; ++p; }
    char digits[24];
    int nd = 0;
    while (p < end && *p != 'e') {
        if (*p != '.') digits[nd++] = *p;
        ++p;
    }
    const int exp10 = (int) std::strtol(p + 1, nullptr, 10);
    const int decpt = exp10 + 1;                 // value = 0.d1d2... x 10^decpt
    if (decpt > -4 && decpt <= 16) {
        if (decpt <= 0) {
            *out++ = '0'; *out++ = '.';
            for (int z = 0; z < -decpt; ++z) *out++ = '0';
            for (int k = 0; k < nd; ++k) *out++ = digits[k];
        } else if (decpt >= nd) {
            for (int k = 0; k < nd; ++k) *out++ = digits[k];
            for (int z = 0; z < decpt - nd; ++z) *out++ = '0';
            *out++ = '.'; *out++ = '0';
        } else {
            for (int k = 0; k < decpt; ++k) *out++ = digits[k];
            *out++ = '.';
            for (int k = decpt; k < nd; ++k) *out++ = digits[k];
        }
    } else {
        *out++ = digits[0];
        if (nd > 1) {
            *out++ = '.';
            for (int k = 1; k < nd; ++k) *out++ = digits[k];
        }
        *out++ = 'e';
        int e = exp10;
        *out++ = e < 0 ? '-' : '+';
        if (e < 0) e = -e;
        char eb[8];
        int ne = 0;
        do { eb[ne++] = (char) ('0' + e % 10); e /= 10; } while (e);
        if (ne < 2) eb[ne++] = '0';
        while (ne) *out++ = eb[--ne];
    }
    return out;
}

int pml_write_float_table(const char *path, const char *header, const char *labels, const int64_t *label_offsets,
                          const double *values, int64_t n_rows, int32_t n_cols) {
    if (!path || !header || !labels || !label_offsets || (!values && n_rows > 0 && n_cols > 0) || n_rows < 0 || n_cols < 0) {
        g_newick_error = "null argument";
        return 1;
    }
    std::FILE *f = std::fopen(path, "wb");
    if (!f) {
        g_newick_error = std::string("cannot open ") + path;
        return 1;
    }
    std::fputs(header, f);
    std::fputc('\n', f);
    std::vector<char> buf;
    buf.reserve(1 << 20);
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t a = label_offsets[r], b = label_offsets[r + 1];
        const size_t need = (size_t) (b - a) + (size_t) n_cols * 28 + 2;
        const size_t at = buf.size();
        buf.resize(at + need);
        char *out = buf.data() + at;
        std::memcpy(out, labels + a, (size_t) (b - a));
        out += b - a;
        for (int c = 0; c < n_cols; ++c) {
            *out++ = '\t';
            out = append_repr(out, values[(size_t) r * n_cols + c]);
        }
        *out++ = '\n';
        buf.resize((size_t) (out - buf.data()));
        if (buf.size() > (1 << 20) - 4096) {
            std::fwrite(buf.data(), 1, buf.size(), f);
            buf.clear();
        }
    }
    if (!buf.empty()) std::fwrite(buf.data(), 1, buf.size(), f);
    const bool ok = std::fclose(f) == 0;
    if (!ok) g_newick_error = std::string("cannot write ") + path;
    return ok ? 0 : 1;
}

}  // extern "C"
