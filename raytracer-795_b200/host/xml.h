// Minimal XML DOM for the scene grammar of the reference.
// Replaces the reference's vendored tinyxml2 for this input surface: elements, attributes, character data,
// comments, <?...?> and <!...> skipped.  API names follow the handful of tinyxml2 calls src/Parser.h makes
// (FirstChildElement / NextSiblingElement / GetText / Query*Text / attributes in document order) so the loader
// reads like the reference parser.  Own code; not derived from tinyxml2.
#pragma once
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

namespace rt795 {

struct XmlAttr { std::string name, value; };

struct XmlElem {
    std::string name;
    std::vector<XmlAttr> attrs;              // document order (the reference scans them with FirstAttribute/Next)
    std::vector<std::unique_ptr<XmlElem>> children;
    XmlElem* parent = nullptr;
    std::string text;                        // text of the FIRST child node if it is character data (tinyxml2 GetText)
    bool has_text = false;

    const XmlElem* FirstChildElement(const char* n = nullptr) const {
        for (auto& c : children) if (!n || c->name == n) return c.get();
        return nullptr;
    }
    const XmlElem* NextSiblingElement(const char* n = nullptr) const {
        if (!parent) return nullptr;
        bool seen = false;
        for (auto& c : parent->children) {
            if (seen && (!n || c->name == n)) return c.get();
            if (c.get() == this) seen = true;
        }
        return nullptr;
    }
    const char* GetText() const { return has_text ? text.c_str() : nullptr; }
    // tinyxml2 XMLUtil::ToInt / ToFloat are sscanf("%d") / sscanf("%f")
    bool QueryIntText(int* v) const { return has_text && sscanf(text.c_str(), "%d", v) == 1; }
    bool QueryFloatText(float* v) const { return has_text && sscanf(text.c_str(), "%f", v) == 1; }
    const XmlAttr* FindAttribute(const char* n) const {
        for (auto& a : attrs) if (a.name == n) return &a;
        return nullptr;
    }
    // The reference looks attributes up by PREFIX: strncmp(attr->Name(), "type", 4) (e.g. src/Parser.h:408-433):
    // first attribute, in document order, whose name starts with `prefix`.
    const XmlAttr* FindAttributePrefix(const char* prefix) const {
        size_t n = strlen(prefix);
        for (auto& a : attrs) if (strncmp(a.name.c_str(), prefix, n) == 0) return &a;
        return nullptr;
    }
    bool QueryIntAttribute(const char* n, int* v) const {
        const XmlAttr* a = FindAttribute(n);
        return a && sscanf(a->value.c_str(), "%d", v) == 1;
    }
    bool QueryBoolAttribute(const char* n, bool* v) const {  // tinyxml2 XMLUtil::ToBool
        const XmlAttr* a = FindAttribute(n);
        if (!a) return false;
        int iv;
        if (sscanf(a->value.c_str(), "%d", &iv) == 1) { *v = iv != 0; return true; }
        const std::string& s = a->value;
        if (s == "true" || s == "True" || s == "TRUE") { *v = true; return true; }
        if (s == "false" || s == "False" || s == "FALSE") { *v = false; return true; }
        return false;
    }
};

class XmlDoc {
public:
    // Returns false and fills err on failure.  root() is the first element node of the file.
    bool LoadFile(const char* path, std::string* err) {
        FILE* f = fopen(path, "rb");
        if (!f) { *err = std::string("cannot open ") + path; return false; }
        std::string s;
        char buf[1 << 16];
        size_t n;
        while ((n = fread(buf, 1, sizeof buf, f)) > 0) s.append(buf, n);
        fclose(f);
        return Parse(s, err);
    }
    bool Parse(const std::string& s, std::string* err) {
        src_ = &s; pos_ = 0; root_.reset();
        XmlElem top;
        if (!ParseContent(&top, err)) return false;
        for (auto& c : top.children) { root_ = std::move(c); break; }
        if (!root_) { *err = "no root element"; return false; }
        root_->parent = nullptr;
        return true;
    }
    const XmlElem* root() const { return root_.get(); }

private:
    const std::string* src_ = nullptr;
    size_t pos_ = 0;
    std::unique_ptr<XmlElem> root_;

    bool starts(const char* lit) const { return src_->compare(pos_, strlen(lit), lit) == 0; }
    static void Decode(std::string& t) {
        if (t.find('&') == std::string::npos) return;
        static const char* ent[][2] = {{"&lt;", "<"}, {"&gt;", ">"}, {"&amp;", "&"}, {"&quot;", "\""}, {"&apos;", "'"}};
        std::string o;
        for (size_t i = 0; i < t.size();) {
            bool hit = false;
            if (t[i] == '&')
                for (auto& e : ent) { size_t l = strlen(e[0]); if (t.compare(i, l, e[0]) == 0) { o += e[1]; i += l; hit = true; break; } }
            if (!hit) o += t[i++];
        }
        t.swap(o);
    }
    bool ParseContent(XmlElem* parent, std::string* err) {
        const std::string& s = *src_;
        bool first_node = true;
        while (pos_ < s.size()) {
            if (s[pos_] != '<') {
                size_t e = s.find('<', pos_);
                if (e == std::string::npos) e = s.size();
                std::string t = s.substr(pos_, e - pos_);
                pos_ = e;
                bool all_ws = t.find_first_not_of(" \t\r\n") == std::string::npos;
                if (!all_ws) {
                    if (first_node && parent->name.size()) { Decode(t); parent->text = t; parent->has_text = true; }
                    first_node = false;
                }
                continue;
            }
            if (starts("<!--")) { size_t e = s.find("-->", pos_); if (e == std::string::npos) { *err = "unterminated comment"; return false; } pos_ = e + 3; first_node = false; continue; }
            if (starts("<?")) { size_t e = s.find("?>", pos_); if (e == std::string::npos) { *err = "unterminated PI"; return false; } pos_ = e + 2; continue; }
            if (starts("<!")) { size_t e = s.find('>', pos_); if (e == std::string::npos) { *err = "unterminated decl"; return false; } pos_ = e + 1; continue; }
            if (starts("</")) {
                size_t e = s.find('>', pos_);
                if (e == std::string::npos) { *err = "unterminated end tag"; return false; }
                std::string nm = s.substr(pos_ + 2, e - pos_ - 2);
                while (!nm.empty() && isspace((unsigned char)nm.back())) nm.pop_back();
                if (nm != parent->name) { *err = "mismatched </" + nm + "> inside <" + parent->name + ">"; return false; }
                pos_ = e + 1;
                return true;
            }
            // start tag
            first_node = false;
            pos_++;
            size_t b = pos_;
            while (pos_ < s.size() && !isspace((unsigned char)s[pos_]) && s[pos_] != '>' && s[pos_] != '/') pos_++;
            std::unique_ptr<XmlElem> el(new XmlElem());
            el->name = s.substr(b, pos_ - b);
            el->parent = parent;
            bool selfclose = false;
            for (;;) {
                while (pos_ < s.size() && isspace((unsigned char)s[pos_])) pos_++;
                if (pos_ >= s.size()) { *err = "unterminated start tag <" + el->name; return false; }
                if (s[pos_] == '>') { pos_++; break; }
                if (s[pos_] == '/') { selfclose = true; pos_++; continue; }
                size_t nb = pos_;
                while (pos_ < s.size() && s[pos_] != '=' && !isspace((unsigned char)s[pos_]) && s[pos_] != '>') pos_++;
                XmlAttr a;
                a.name = s.substr(nb, pos_ - nb);
                while (pos_ < s.size() && isspace((unsigned char)s[pos_])) pos_++;
                if (pos_ < s.size() && s[pos_] == '=') {
                    pos_++;
                    while (pos_ < s.size() && isspace((unsigned char)s[pos_])) pos_++;
                    char q = s[pos_];
                    if (q != '"' && q != '\'') { *err = "attribute value not quoted in <" + el->name + ">"; return false; }
                    size_t e = s.find(q, pos_ + 1);
                    if (e == std::string::npos) { *err = "unterminated attribute value"; return false; }
                    a.value = s.substr(pos_ + 1, e - pos_ - 1);
                    Decode(a.value);
                    pos_ = e + 1;
                }
                el->attrs.push_back(a);
            }
            XmlElem* raw = el.get();
            parent->children.push_back(std::move(el));
            if (!selfclose) {
                // text counts as GetText() only if it is the first child node of the element
                if (!ParseContent(raw, err)) return false;
            }
        }
        if (parent->name.size()) { *err = "unexpected end of file inside <" + parent->name + ">"; return false; }
        return true;
    }
};

}  // namespace rt795
