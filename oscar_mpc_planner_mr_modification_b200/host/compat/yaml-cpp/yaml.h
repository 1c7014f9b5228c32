// Minimal stand-in for yaml-cpp (absent from this image): exactly the subset the Solver / State classes
// use -- LoadFile, node[key], node[index], as<T>(), IsDefined(), ordered map iteration.  Block-style
// YAML only (maps, "- " sequences, scalars, comments, quoted strings).  In a ROS workspace the real
// yaml-cpp is found first and this header is not on the include path.
#pragma once
#include <fstream>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace YAML {
class Node;
struct NodeData {
    enum Kind { Undefined, Scalar, Map, Seq } kind = Undefined;
    std::string scalar;
    std::vector<std::pair<Node, Node>> map;
    std::vector<Node> seq;
};

class Node {
public:
    Node() : d_(std::make_shared<NodeData>()) {}
    explicit Node(const std::string& s) : d_(std::make_shared<NodeData>()) { d_->kind = NodeData::Scalar; d_->scalar = s; }
    bool IsDefined() const { return d_->kind != NodeData::Undefined; }
    bool IsMap() const { return d_->kind == NodeData::Map; }
    bool IsSequence() const { return d_->kind == NodeData::Seq; }
    size_t size() const { return IsMap() ? d_->map.size() : d_->seq.size(); }
    const Node operator[](const std::string& key) const
    {
        for (auto& kv : d_->map)
            if (kv.first.d_->scalar == key) return kv.second;
        return Node();
    }
    const Node operator[](const char* key) const { return (*this)[std::string(key)]; }
    const Node operator[](int i) const { return (i >= 0 && i < (int)d_->seq.size()) ? d_->seq[i] : Node(); }
    template <typename T> T as() const;
    typedef std::vector<std::pair<Node, Node>>::const_iterator const_iterator;
    const_iterator begin() const { return d_->map.begin(); }
    const_iterator end() const { return d_->map.end(); }
    NodeData& data() { return *d_; }
    const NodeData& data() const { return *d_; }

private:
    std::shared_ptr<NodeData> d_;
};
typedef Node::const_iterator const_iterator;

template <> inline std::string Node::as<std::string>() const
{
    if (d_->kind != NodeData::Scalar) throw std::runtime_error("YAML: not a scalar");
    return d_->scalar;
}
template <> inline double Node::as<double>() const { return std::stod(as<std::string>()); }
template <> inline int Node::as<int>() const { return (int)std::stol(as<std::string>()); }
template <> inline unsigned int Node::as<unsigned int>() const { return (unsigned int)std::stoul(as<std::string>()); }
template <> inline bool Node::as<bool>() const
{
    const std::string s = as<std::string>();
    return s == "true" || s == "True" || s == "yes" || s == "1";
}

namespace detail {
struct Line { int indent; std::string text; };
inline std::string strip(const std::string& s)
{
    size_t a = s.find_first_not_of(" \t\r"), b = s.find_last_not_of(" \t\r");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
inline std::string unquote(std::string s)
{
    s = strip(s);
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}
inline std::string drop_comment(const std::string& s)
{
    char q = 0;
    for (size_t i = 0; i < s.size(); i++) {
        if (q) { if (s[i] == q) q = 0; }
        else if (s[i] == '"' || s[i] == '\'') q = s[i];
        else if (s[i] == '#' && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}
inline Node parse_block(const std::vector<Line>& L, size_t& i, int indent);
inline Node parse_value(const std::vector<Line>& L, size_t& i, int key_indent, const std::string& rest)
{
    if (!rest.empty()) return Node(unquote(rest));
    if (i < L.size() && (L[i].indent > key_indent || (L[i].indent == key_indent && L[i].text.compare(0, 1, "-") == 0)))
        return parse_block(L, i, L[i].indent);
    return Node(std::string());
}
inline Node parse_block(const std::vector<Line>& L, size_t& i, int indent)
{
    Node n;
    const bool seq = L[i].text[0] == '-' && (L[i].text.size() == 1 || L[i].text[1] == ' ');
    n.data().kind = seq ? NodeData::Seq : NodeData::Map;
    while (i < L.size() && L[i].indent == indent) {
        const std::string t = L[i].text;
        if (seq) {
            if (!(t[0] == '-' && (t.size() == 1 || t[1] == ' '))) break;
            i++;
            n.data().seq.push_back(parse_value(L, i, indent, strip(t.substr(1))));
        } else {
            size_t c = t.find(':');
            if (c == std::string::npos) throw std::runtime_error("YAML: expected 'key:' in line '" + t + "'");
            i++;
            Node key(unquote(t.substr(0, c)));
            n.data().map.emplace_back(key, parse_value(L, i, indent, strip(t.substr(c + 1))));
        }
    }
    return n;
}
}  // namespace detail

inline Node Load(std::istream& in)
{
    std::vector<detail::Line> L;
    std::string s;
    while (std::getline(in, s)) {
        s = detail::drop_comment(s);
        const std::string t = detail::strip(s);
        if (t.empty() || t == "---") continue;
        L.push_back({(int)s.find_first_not_of(" \t"), t});
    }
    if (L.empty()) return Node();
    size_t i = 0;
    return detail::parse_block(L, i, L[0].indent);
}
inline Node LoadFile(const std::string& path)
{
    std::ifstream f(path);
    if (!f) throw std::runtime_error("YAML: cannot open " + path);
    return Load(f);
}
}  // namespace YAML
