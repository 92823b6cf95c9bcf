// Minimal XML DOM reader for URDF / SDF model descriptions (elements, attributes, text, comments,
// declarations, CDATA). Replaces the tinyxml2 layer underneath sdformat for the subset the loader needs
// (reference: cpp/scenario/gazebo/src/helpers.cpp:48-88 getSdfRootFromFile/String).
#pragma once

#include <cctype>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace b2 {

struct XmlNode {
    std::string tag;
    std::vector<std::pair<std::string, std::string>> attrs;
    std::vector<std::unique_ptr<XmlNode>> children;
    std::string text;

    const char* attr(const char* name) const
    {
        for (auto& a : attrs)
            if (a.first == name) return a.second.c_str();
        return nullptr;
    }
    const XmlNode* child(const char* name) const
    {
        for (auto& c : children)
            if (c->tag == name) return c.get();
        return nullptr;
    }
    std::vector<const XmlNode*> all(const char* name) const
    {
        std::vector<const XmlNode*> out;
        for (auto& c : children)
            if (c->tag == name) out.push_back(c.get());
        return out;
    }
};

class XmlParser {
public:
    XmlParser(const char* data, size_t len) : s_(data), n_(len) {}

    std::unique_ptr<XmlNode> parse()
    {
        skip_misc();
        auto root = element();
        if (!root) throw std::runtime_error("XML: no root element");
        return root;
    }

private:
    const char* s_;
    size_t n_;
    size_t i_ = 0;

    bool starts(const char* lit) const
    {
        size_t l = strlen(lit);
        return i_ + l <= n_ && memcmp(s_ + i_, lit, l) == 0;
    }
    void skip_ws()
    {
        while (i_ < n_ && isspace((unsigned char)s_[i_])) ++i_;
    }
    void skip_until(const char* lit)
    {
        size_t l = strlen(lit);
        while (i_ + l <= n_ && memcmp(s_ + i_, lit, l) != 0) ++i_;
        if (i_ + l > n_) throw std::runtime_error(std::string("XML: unterminated construct, expected ") + lit);
        i_ += l;
    }
    void skip_misc()
    {
        for (;;) {
            skip_ws();
            if (starts("<?")) skip_until("?>");
            else if (starts("<!--")) skip_until("-->");
            else if (starts("<!DOCTYPE")) skip_until(">");
            else return;
        }
    }
    std::string name()
    {
        size_t b = i_;
        while (i_ < n_ && (isalnum((unsigned char)s_[i_]) || s_[i_] == '_' || s_[i_] == ':' || s_[i_] == '-' ||
                           s_[i_] == '.'))
            ++i_;
        if (b == i_) throw std::runtime_error("XML: expected a name at offset " + std::to_string(i_));
        return std::string(s_ + b, i_ - b);
    }
    static std::string unescape(const std::string& in)
    {
        std::string out;
        for (size_t k = 0; k < in.size(); ++k) {
            if (in[k] != '&') { out.push_back(in[k]); continue; }
            if (!in.compare(k, 4, "&lt;")) { out.push_back('<'); k += 3; }
            else if (!in.compare(k, 4, "&gt;")) { out.push_back('>'); k += 3; }
            else if (!in.compare(k, 5, "&amp;")) { out.push_back('&'); k += 4; }
            else if (!in.compare(k, 6, "&quot;")) { out.push_back('"'); k += 5; }
            else if (!in.compare(k, 6, "&apos;")) { out.push_back('\''); k += 5; }
            else out.push_back('&');
        }
        return out;
    }
    std::unique_ptr<XmlNode> element()
    {
        if (i_ >= n_ || s_[i_] != '<') return nullptr;
        ++i_;
        auto node = std::make_unique<XmlNode>();
        node->tag = name();
        for (;;) {
            skip_ws();
            if (i_ >= n_) throw std::runtime_error("XML: unexpected end inside <" + node->tag + ">");
            if (starts("/>")) { i_ += 2; return node; }
            if (s_[i_] == '>') { ++i_; break; }
            std::string key = name();
            skip_ws();
            if (i_ >= n_ || s_[i_] != '=') throw std::runtime_error("XML: expected '=' after attribute " + key);
            ++i_;
            skip_ws();
            if (i_ >= n_ || (s_[i_] != '"' && s_[i_] != '\'')) throw std::runtime_error("XML: unquoted attribute " + key);
            char quote = s_[i_++];
            size_t b = i_;
            while (i_ < n_ && s_[i_] != quote) ++i_;
            if (i_ >= n_) throw std::runtime_error("XML: unterminated attribute " + key);
            node->attrs.emplace_back(key, unescape(std::string(s_ + b, i_ - b)));
            ++i_;
        }
        // content
        for (;;) {
            if (i_ >= n_) throw std::runtime_error("XML: missing </" + node->tag + ">");
            if (starts("<!--")) { skip_until("-->"); continue; }
            if (starts("<![CDATA[")) {
                i_ += 9;
                size_t b = i_;
                skip_until("]]>");
                node->text.append(s_ + b, i_ - 3 - b);
                continue;
            }
            if (starts("<?")) { skip_until("?>"); continue; }
            if (starts("</")) {
                i_ += 2;
                std::string closing = name();
                if (closing != node->tag) throw std::runtime_error("XML: </" + closing + "> closes <" + node->tag + ">");
                skip_ws();
                if (i_ >= n_ || s_[i_] != '>') throw std::runtime_error("XML: malformed closing tag " + closing);
                ++i_;
                break;
            }
            if (s_[i_] == '<') {
                node->children.push_back(element());
                continue;
            }
            size_t b = i_;
            while (i_ < n_ && s_[i_] != '<') ++i_;
            node->text += unescape(std::string(s_ + b, i_ - b));
        }
        // trim text
        size_t b = 0, e = node->text.size();
        while (b < e && isspace((unsigned char)node->text[b])) ++b;
        while (e > b && isspace((unsigned char)node->text[e - 1])) --e;
        node->text = node->text.substr(b, e - b);
        return node;
    }
};

}  // namespace b2
