// Minimal self-written stand-in for the yaml-cpp calls the PBTE host layer makes
// (YAML::LoadFile, Node::operator[], Node::as<T>, operator bool).  Supports block maps, block
// sequences, flow sequences "[a, b]", scalars and '#' comments — the subset config.yaml / si.yaml use.
// yaml-cpp itself is not installed in the build container; deployments link the real library.
#pragma once

#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace YAML
{
class Exception : public std::runtime_error
{
public:
    explicit Exception(const std::string &m) : std::runtime_error(m) {}
};
class BadFile : public Exception
{
public:
    explicit BadFile(const std::string &f) : Exception("bad file: " + f) {}
};
class BadConversion : public Exception
{
public:
    explicit BadConversion(const std::string &v) : Exception("bad conversion: '" + v + "'") {}
};

class Node
{
public:
    enum Kind { Undefined, Null, Scalar, Sequence, Map };
    Node() = default;
    explicit operator bool() const { return d_ && d_->kind != Undefined; }
    bool operator!() const { return !static_cast<bool>(*this); }
    bool IsDefined() const { return static_cast<bool>(*this); }
    bool IsNull() const { return d_ && d_->kind == Null; }
    bool IsScalar() const { return d_ && d_->kind == Scalar; }
    bool IsSequence() const { return d_ && d_->kind == Sequence; }
    bool IsMap() const { return d_ && d_->kind == Map; }
    std::size_t size() const { return !d_ ? 0 : d_->kind == Sequence ? d_->seq.size() : d_->kind == Map ? d_->keys.size() : 0; }

    const Node operator[](const std::string &key) const
    {
        if (d_ && d_->kind == Map)
        {
            auto it = d_->map.find(key);
            if (it != d_->map.end()) return it->second;
        }
        return Node();
    }
    const Node operator[](const char *key) const { return (*this)[std::string(key)]; }
    const Node operator[](int i) const
    {
        if (d_ && d_->kind == Sequence && i >= 0 && i < static_cast<int>(d_->seq.size())) return d_->seq[i];
        return Node();
    }
    const Node operator[](std::size_t i) const { return (*this)[static_cast<int>(i)]; }

    const std::string &Scalar_() const
    {
        if (!d_ || d_->kind != Scalar) throw BadConversion("<non-scalar>");
        return d_->scalar;
    }
    template <class T>
    T as() const
    {
        return Convert<T>::get(*this);
    }
    template <class T>
    T as(const T &fallback) const
    {
        try
        {
            return Convert<T>::get(*this);
        }
        catch (const Exception &)
        {
            return fallback;
        }
    }

    // construction (used by the parser)
    static Node MakeScalar(const std::string &s)
    {
        Node n;
        n.d_ = std::make_shared<Data>();
        if (s.empty() || s == "~" || s == "null") n.d_->kind = Null;
        else
        {
            n.d_->kind = Scalar;
            n.d_->scalar = s;
        }
        return n;
    }
    static Node MakeSeq()
    {
        Node n;
        n.d_ = std::make_shared<Data>();
        n.d_->kind = Sequence;
        return n;
    }
    static Node MakeMap()
    {
        Node n;
        n.d_ = std::make_shared<Data>();
        n.d_->kind = Map;
        return n;
    }
    void push_back(const Node &c) { d_->seq.push_back(c); }
    void set(const std::string &k, const Node &c)
    {
        if (!d_->map.count(k)) d_->keys.push_back(k);
        d_->map[k] = c;
    }

private:
    template <class T, class = void>
    struct Convert;
    struct Data
    {
        Kind kind = Undefined;
        std::string scalar;
        std::vector<Node> seq;
        std::map<std::string, Node> map;
        std::vector<std::string> keys;
    };
    std::shared_ptr<Data> d_;
};

template <>
struct Node::Convert<std::string>
{
    static std::string get(const Node &n) { return n.Scalar_(); }
};
template <>
struct Node::Convert<double>
{
    static double get(const Node &n)
    {
        const std::string &s = n.Scalar_();
        try
        {
            std::size_t pos = 0;
            const double v = std::stod(s, &pos);
            if (pos != s.size()) throw BadConversion(s);
            return v;
        }
        catch (const std::logic_error &)
        {
            throw BadConversion(s);
        }
    }
};
template <>
struct Node::Convert<int>
{
    static int get(const Node &n)
    {
        const std::string &s = n.Scalar_();
        try
        {
            std::size_t pos = 0;
            const int v = std::stoi(s, &pos);
            if (pos != s.size()) throw BadConversion(s);
            return v;
        }
        catch (const std::logic_error &)
        {
            throw BadConversion(s);
        }
    }
};
template <>
struct Node::Convert<bool>
{
    static bool get(const Node &n)
    {
        const std::string &s = n.Scalar_();
        if (s == "true" || s == "True" || s == "yes" || s == "on") return true;
        if (s == "false" || s == "False" || s == "no" || s == "off") return false;
        throw BadConversion(s);
    }
};
template <class T>
struct Node::Convert<std::vector<T>>
{
    static std::vector<T> get(const Node &n)
    {
        if (!n.IsSequence()) throw BadConversion("<non-sequence>");
        std::vector<T> out;
        for (std::size_t i = 0; i < n.size(); ++i) out.push_back(n[i].template as<T>());
        return out;
    }
};

namespace detail
{
struct Line
{
    int indent;
    std::string text;
};

inline std::string Trim(const std::string &s)
{
    const auto a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    const auto b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

inline std::string StripComment(const std::string &s)
{
    bool sq = false, dq = false;
    for (std::size_t i = 0; i < s.size(); ++i)
    {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == '#' && !sq && !dq && (i == 0 || s[i - 1] == ' ' || s[i - 1] == '\t')) return s.substr(0, i);
    }
    return s;
}

inline std::string Unquote(const std::string &s)
{
    if (s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) return s.substr(1, s.size() - 2);
    return s;
}

inline Node ParseInline(const std::string &raw)
{
    const std::string s = Trim(raw);
    if (!s.empty() && s.front() == '[')
    {
        if (s.back() != ']') throw Exception("unterminated flow sequence: " + s);
        Node seq = Node::MakeSeq();
        std::string item;
        int depth = 0;
        for (std::size_t i = 1; i + 1 < s.size(); ++i)
        {
            const char c = s[i];
            if (c == '[') ++depth;
            if (c == ']') --depth;
            if (c == ',' && depth == 0)
            {
                seq.push_back(ParseInline(item));
                item.clear();
            }
            else item += c;
        }
        if (!Trim(item).empty()) seq.push_back(ParseInline(item));
        return seq;
    }
    return Node::MakeScalar(Unquote(s));
}

// position of the "key:" separator (a ':' followed by space/end, outside quotes), or npos
inline std::size_t KeySep(const std::string &s)
{
    bool sq = false, dq = false;
    for (std::size_t i = 0; i < s.size(); ++i)
    {
        const char c = s[i];
        if (c == '\'' && !dq) sq = !sq;
        else if (c == '"' && !sq) dq = !dq;
        else if (c == ':' && !sq && !dq && (i + 1 == s.size() || s[i + 1] == ' ' || s[i + 1] == '\t')) return i;
    }
    return std::string::npos;
}

inline Node ParseBlock(std::vector<Line> &lines, std::size_t &pos, int indent);

inline Node ParseMapFrom(std::vector<Line> &lines, std::size_t &pos, int indent)
{
    Node map = Node::MakeMap();
    while (pos < lines.size() && lines[pos].indent == indent && lines[pos].text.rfind("- ", 0) != 0 && lines[pos].text != "-")
    {
        const std::string t = lines[pos].text;
        const std::size_t sep = KeySep(t);
        if (sep == std::string::npos) throw Exception("expected 'key: value' at: " + t);
        const std::string key = Unquote(Trim(t.substr(0, sep)));
        const std::string rest = Trim(t.substr(sep + 1));
        ++pos;
        if (!rest.empty()) map.set(key, ParseInline(rest));
        else if (pos < lines.size() && (lines[pos].indent > indent || (lines[pos].indent == indent && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-"))))
            map.set(key, ParseBlock(lines, pos, lines[pos].indent));
        else map.set(key, Node::MakeScalar(""));
    }
    return map;
}

inline Node ParseBlock(std::vector<Line> &lines, std::size_t &pos, int indent)
{
    if (pos >= lines.size()) return Node::MakeScalar("");
    const std::string &first = lines[pos].text;
    if (first.rfind("- ", 0) == 0 || first == "-")
    {
        Node seq = Node::MakeSeq();
        while (pos < lines.size() && lines[pos].indent == indent && (lines[pos].text.rfind("- ", 0) == 0 || lines[pos].text == "-"))
        {
            const std::string rest = Trim(lines[pos].text.substr(1));
            if (rest.empty())
            {
                ++pos;
                seq.push_back(pos < lines.size() && lines[pos].indent > indent ? ParseBlock(lines, pos, lines[pos].indent) : Node::MakeScalar(""));
            }
            else if (KeySep(rest) != std::string::npos && rest.front() != '[')
            {
                // "- key: value" opens a map whose keys are indented past the dash
                const int sub = indent + static_cast<int>(lines[pos].text.size() - rest.size());
                lines[pos].indent = sub;
                lines[pos].text = rest;
                seq.push_back(ParseMapFrom(lines, pos, sub));
            }
            else
            {
                seq.push_back(ParseInline(rest));
                ++pos;
            }
        }
        return seq;
    }
    if (KeySep(first) != std::string::npos) return ParseMapFrom(lines, pos, indent);
    Node n = ParseInline(first);
    ++pos;
    return n;
}
}  // namespace detail

inline Node Load(const std::string &text)
{
    std::vector<detail::Line> lines;
    std::istringstream in(text);
    std::string raw;
    while (std::getline(in, raw))
    {
        const std::string s = detail::StripComment(raw);
        const std::string t = detail::Trim(s);
        if (t.empty() || t == "---") continue;
        int indent = 0;
        while (indent < static_cast<int>(s.size()) && s[indent] == ' ') ++indent;
        lines.push_back({indent, t});
    }
    if (lines.empty()) return Node::MakeScalar("");
    std::size_t pos = 0;
    return detail::ParseBlock(lines, pos, lines[0].indent);
}

inline Node LoadFile(const std::string &path)
{
    std::ifstream in(path);
    if (!in) throw BadFile(path);
    std::stringstream ss;
    ss << in.rdbuf();
    return Load(ss.str());
}
}  // namespace YAML
