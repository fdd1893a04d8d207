// yaml_lite.hpp -- the subset of YAML the reference's scene / animation files use
// (Data.Yaml.decodeFileEither, app/Main.hs:85): block mappings by indentation, block
// sequences ("- "), flow sequences [a, b], flow mappings {k: v}, plain / quoted scalars,
// '#' comments.  No anchors, tags, multi-line scalars or documents.
#pragma once

#include <cctype>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace yamllite {

struct Node;
using NodeP = std::shared_ptr<Node>;

struct Node {
    enum Kind { Null, Scalar, Seq, Map } kind = Null;
    std::string scalar;
    std::vector<NodeP> seq;
    std::vector<std::pair<std::string, NodeP>> map;  // insertion order

    const Node *get(const std::string &key) const
    {
        if (kind != Map) return nullptr;
        for (auto &kv : map)
            if (kv.first == key) return kv.second.get();
        return nullptr;
    }
    bool is_null() const { return kind == Null || (kind == Scalar && (scalar == "~" || scalar == "null" || scalar.empty())); }
    double as_double(const std::string &what) const
    {
        if (kind != Scalar) throw std::runtime_error(what + ": expected a number");
        char *end = nullptr;
        const double v = std::strtod(scalar.c_str(), &end);
        if (end == scalar.c_str() || *end != '\0') throw std::runtime_error(what + ": expected a number, got '" + scalar + "'");
        return v;
    }
    long as_int(const std::string &what) const
    {
        const double v = as_double(what);
        if (v != (double)(long)v) throw std::runtime_error(what + ": expected an integer");
        return (long)v;
    }
    bool as_bool(const std::string &what) const
    {
        if (kind == Scalar) {
            if (scalar == "true" || scalar == "True" || scalar == "yes" || scalar == "on") return true;
            if (scalar == "false" || scalar == "False" || scalar == "no" || scalar == "off") return false;
        }
        throw std::runtime_error(what + ": expected a boolean");
    }
};

namespace detail {

struct Line { int indent; std::string text; int no; };

inline std::string strip_comment(const std::string &s)
{
    bool sq = false, dq = false;
    for (size_t i = 0; i < s.size(); i++) {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == '#' && !sq && !dq && (i == 0 || std::isspace((unsigned char)s[i - 1]))) return s.substr(0, i);
    }
    return s;
}
inline std::string trim(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && std::isspace((unsigned char)s[a])) a++;
    while (b > a && std::isspace((unsigned char)s[b - 1])) b--;
    return s.substr(a, b - a);
}
inline std::string unquote(const std::string &s)
{
    if (s.size() >= 2 && ((s.front() == '\'' && s.back() == '\'') || (s.front() == '"' && s.back() == '"')))
        return s.substr(1, s.size() - 2);
    return s;
}

// ---- flow syntax: [ ... ] and { ... }
struct Flow {
    const std::string &s; size_t p = 0; int line;
    Flow(const std::string &str, int ln) : s(str), line(ln) {}
    void ws() { while (p < s.size() && std::isspace((unsigned char)s[p])) p++; }
    [[noreturn]] void fail(const std::string &m) { throw std::runtime_error("line " + std::to_string(line) + ": " + m); }
    NodeP value()
    {
        ws();
        if (p >= s.size()) return std::make_shared<Node>();
        if (s[p] == '[') return seq();
        if (s[p] == '{') return map();
        return scalar(",]}");
    }
    NodeP scalar(const char *stops)
    {
        ws();
        auto n = std::make_shared<Node>();
        n->kind = Node::Scalar;
        if (p < s.size() && (s[p] == '\'' || s[p] == '"')) {
            const char q = s[p++];
            const size_t b = p;
            while (p < s.size() && s[p] != q) p++;
            if (p >= s.size()) fail("unterminated quoted scalar");
            n->scalar = s.substr(b, p - b);
            p++;
            return n;
        }
        const size_t b = p;
        while (p < s.size() && !std::strchr(stops, s[p])) p++;
        n->scalar = trim(s.substr(b, p - b));
        return n;
    }
    NodeP seq()
    {
        auto n = std::make_shared<Node>();
        n->kind = Node::Seq;
        p++;  // [
        for (;;) {
            ws();
            if (p >= s.size()) fail("unterminated flow sequence");
            if (s[p] == ']') { p++; break; }
            n->seq.push_back(value());
            ws();
            if (p < s.size() && s[p] == ',') p++;
        }
        return n;
    }
    NodeP map()
    {
        auto n = std::make_shared<Node>();
        n->kind = Node::Map;
        p++;  // {
        for (;;) {
            ws();
            if (p >= s.size()) fail("unterminated flow mapping");
            if (s[p] == '}') { p++; break; }
            NodeP k = scalar(":,}");
            ws();
            if (p >= s.size() || s[p] != ':') fail("expected ':' in flow mapping");
            p++;
            n->map.emplace_back(k->scalar, value());
            ws();
            if (p < s.size() && s[p] == ',') p++;
        }
        return n;
    }
};

inline NodeP inline_value(const std::string &text, int line)
{
    const std::string t = trim(text);
    if (t.empty()) return std::make_shared<Node>();
    if (t[0] == '[' || t[0] == '{') {
        Flow f(t, line);
        NodeP n = f.value();
        f.ws();
        if (f.p != t.size()) f.fail("trailing characters after flow collection");
        return n;
    }
    auto n = std::make_shared<Node>();
    n->kind = Node::Scalar;
    n->scalar = unquote(t);
    return n;
}

// position of the ':' that separates key and value of a block mapping entry, or npos
inline size_t key_colon(const std::string &t)
{
    bool sq = false, dq = false;
    for (size_t i = 0; i < t.size(); i++) {
        const char c = t[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if ((c == '[' || c == '{') && !sq && !dq) return std::string::npos;
        else if (c == ':' && !sq && !dq && (i + 1 == t.size() || std::isspace((unsigned char)t[i + 1]))) return i;
    }
    return std::string::npos;
}

struct Parser {
    std::vector<Line> lines;
    size_t i = 0;
    [[noreturn]] void fail(const Line &l, const std::string &m) { throw std::runtime_error("line " + std::to_string(l.no) + ": " + m); }

    NodeP block(int indent)
    {
        if (i >= lines.size() || lines[i].indent < indent) return std::make_shared<Node>();
        const int ind = lines[i].indent;
        if (lines[i].text.rfind("- ", 0) == 0 || lines[i].text == "-") return seq(ind);
        if (key_colon(lines[i].text) != std::string::npos) return map(ind);
        NodeP n = inline_value(lines[i].text, lines[i].no);
        i++;
        return n;
    }
    NodeP map(int ind)
    {
        auto n = std::make_shared<Node>();
        n->kind = Node::Map;
        while (i < lines.size() && lines[i].indent == ind) {
            const Line &l = lines[i];
            const size_t c = key_colon(l.text);
            if (c == std::string::npos) fail(l, "expected 'key: value'");
            const std::string key = unquote(trim(l.text.substr(0, c)));
            const std::string rest = trim(l.text.substr(c + 1));
            i++;
            if (!rest.empty()) n->map.emplace_back(key, inline_value(rest, l.no));
            else if (i < lines.size() && lines[i].indent > ind) n->map.emplace_back(key, block(lines[i].indent));
            else if (i < lines.size() && lines[i].indent == ind && lines[i].text.rfind("- ", 0) == 0)
                n->map.emplace_back(key, seq(ind));  // "key:" followed by a sequence at the same indent
            else n->map.emplace_back(key, std::make_shared<Node>());
        }
        if (i < lines.size() && lines[i].indent > ind) fail(lines[i], "bad indentation");
        return n;
    }
    NodeP seq(int ind)
    {
        auto n = std::make_shared<Node>();
        n->kind = Node::Seq;
        while (i < lines.size() && lines[i].indent == ind && (lines[i].text.rfind("- ", 0) == 0 || lines[i].text == "-")) {
            Line &l = lines[i];
            const std::string rest = l.text.size() > 2 ? trim(l.text.substr(2)) : "";
            if (rest.empty()) {
                i++;
                n->seq.push_back(block(ind + 1));
            } else if (key_colon(rest) != std::string::npos && rest[0] != '{' && rest[0] != '[') {
                // "- key: value" starts a mapping whose entries are indented past the dash
                const int child = ind + 2;
                l.text = rest;
                l.indent = child;
                n->seq.push_back(map(child));
            } else {
                n->seq.push_back(inline_value(rest, l.no));
                i++;
            }
        }
        return n;
    }
};

}  // namespace detail

// bracket depth of flow collections left open at the end of `t` (quotes respected)
inline int flow_depth(const std::string &t, int depth)
{
    bool sq = false, dq = false;
    for (char c : t) {
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (!sq && !dq) {
            if (c == '[' || c == '{') depth++;
            else if (c == ']' || c == '}') depth--;
        }
    }
    return depth;
}

inline NodeP parse(const std::string &text)
{
    detail::Parser p;
    size_t pos = 0;
    int no = 0, open = 0;
    while (pos <= text.size()) {
        size_t e = text.find('\n', pos);
        if (e == std::string::npos) e = text.size();
        std::string raw = text.substr(pos, e - pos);
        pos = e + 1;
        no++;
        if (!raw.empty() && raw.back() == '\r') raw.pop_back();
        raw = detail::strip_comment(raw);
        int ind = 0;
        while (ind < (int)raw.size() && raw[ind] == ' ') ind++;
        const std::string t = detail::trim(raw);
        if (open > 0) {
            // continuation of a flow collection that spans lines (PyYAML wraps long ones)
            if (!t.empty()) p.lines.back().text += " " + t;
            open = flow_depth(t, open);
            continue;
        }
        if (ind < (int)raw.size() && raw[ind] == '\t') throw std::runtime_error("line " + std::to_string(no) + ": tab in indentation");
        if (t.empty() || t == "---") continue;
        p.lines.push_back({ ind, t, no });
        open = flow_depth(t, 0);
        if (open < 0) open = 0;
    }
    if (open > 0) throw std::runtime_error("unterminated flow collection at end of input");
    if (p.lines.empty()) return std::make_shared<Node>();
    NodeP root = p.block(0);
    if (p.i < p.lines.size()) p.fail(p.lines[p.i], "unexpected content");
    return root;
}

}  // namespace yamllite
