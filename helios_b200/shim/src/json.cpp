// json.cpp — recursive-descent parser for utility/json.h (RFC 8259 grammar).
#include <utility/json.h>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace helios
{
namespace json
{
namespace
{
    struct Parser
    {
        const std::string& s;
        size_t             p = 0;
        explicit Parser(const std::string& text) : s(text) {}
        [[noreturn]] void fail(const char* what) const { throw std::runtime_error(std::string("json: ") + what + " at byte " + std::to_string(p)); }
        void              skip()
        {
            while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++;
        }
        bool literal(const char* word)
        {
            size_t n = 0;
            while (word[n]) n++;
            if (s.compare(p, n, word) != 0) return false;
            p += n;
            return true;
        }
        static void append_utf8(std::string& out, uint32_t cp)
        {
            if (cp < 0x80)
                out.push_back((char)cp);
            else if (cp < 0x800)
                out.push_back((char)(0xC0 | (cp >> 6))), out.push_back((char)(0x80 | (cp & 0x3F)));
            else if (cp < 0x10000)
                out.push_back((char)(0xE0 | (cp >> 12))), out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))), out.push_back((char)(0x80 | (cp & 0x3F)));
            else
                out.push_back((char)(0xF0 | (cp >> 18))), out.push_back((char)(0x80 | ((cp >> 12) & 0x3F))), out.push_back((char)(0x80 | ((cp >> 6) & 0x3F))), out.push_back((char)(0x80 | (cp & 0x3F)));
        }
        uint32_t hex4()
        {
            if (p + 4 > s.size()) fail("truncated \\u escape");
            uint32_t v = 0;
            for (int i = 0; i < 4; i++)
            {
                const char c = s[p++];
                v <<= 4;
                if (c >= '0' && c <= '9')
                    v |= (uint32_t)(c - '0');
                else if (c >= 'a' && c <= 'f')
                    v |= (uint32_t)(c - 'a' + 10);
                else if (c >= 'A' && c <= 'F')
                    v |= (uint32_t)(c - 'A' + 10);
                else
                    fail("bad \\u escape");
            }
            return v;
        }
        std::string string()
        {
            std::string out;
            p++; // opening quote
            for (;;)
            {
                if (p >= s.size()) fail("unterminated string");
                const char c = s[p++];
                if (c == '"') return out;
                if (c != '\\')
                {
                    out.push_back(c);
                    continue;
                }
                if (p >= s.size()) fail("unterminated escape");
                const char e = s[p++];
                switch (e)
                {
                    case '"': out.push_back('"'); break;
                    case '\\': out.push_back('\\'); break;
                    case '/': out.push_back('/'); break;
                    case 'b': out.push_back('\b'); break;
                    case 'f': out.push_back('\f'); break;
                    case 'n': out.push_back('\n'); break;
                    case 'r': out.push_back('\r'); break;
                    case 't': out.push_back('\t'); break;
                    case 'u':
                    {
                        uint32_t cp = hex4();
                        if (cp >= 0xD800 && cp < 0xDC00 && s.compare(p, 2, "\\u") == 0)
                        {
                            p += 2;
                            const uint32_t lo = hex4();
                            cp                = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
                        }
                        append_utf8(out, cp);
                        break;
                    }
                    default: fail("unknown escape");
                }
            }
        }
        Value value(int depth)
        {
            if (depth > 256) fail("nesting too deep");
            skip();
            if (p >= s.size()) fail("unexpected end");
            Value      v;
            const char c = s[p];
            if (c == '{')
            {
                v.type = Value::Object;
                p++;
                skip();
                if (p < s.size() && s[p] == '}') return p++, v;
                for (;;)
                {
                    skip();
                    if (p >= s.size() || s[p] != '"') fail("expected a key");
                    std::string key = string();
                    skip();
                    if (p >= s.size() || s[p] != ':') fail("expected ':'");
                    p++;
                    v.object.emplace_back(std::move(key), value(depth + 1));
                    skip();
                    if (p < s.size() && s[p] == ',')
                    {
                        p++;
                        continue;
                    }
                    if (p < s.size() && s[p] == '}') return p++, v;
                    fail("expected ',' or '}'");
                }
            }
            if (c == '[')
            {
                v.type = Value::Array;
                p++;
                skip();
                if (p < s.size() && s[p] == ']') return p++, v;
                for (;;)
                {
                    v.array.push_back(value(depth + 1));
                    skip();
                    if (p < s.size() && s[p] == ',')
                    {
                        p++;
                        continue;
                    }
                    if (p < s.size() && s[p] == ']') return p++, v;
                    fail("expected ',' or ']'");
                }
            }
            if (c == '"')
            {
                v.type   = Value::String;
                v.string = string();
                return v;
            }
            if (literal("true")) return v.type = Value::Bool, v.boolean = true, v;
            if (literal("false")) return v.type = Value::Bool, v.boolean = false, v;
            if (literal("null")) return v;
            if (c == '-' || (c >= '0' && c <= '9'))
            {
                const char* begin = s.c_str() + p;
                char*       end   = nullptr;
                v.number          = std::strtod(begin, &end);
                if (end == begin) fail("bad number");
                p += (size_t)(end - begin);
                v.type = Value::Number;
                return v;
            }
            fail("unexpected character");
        }
    };
} // namespace

const Value* Value::find(const std::string& key) const
{
    if (type != Object) return nullptr;
    for (size_t i = object.size(); i-- > 0;)
        if (object[i].first == key) return &object[i].second;
    return nullptr;
}
float Value::get_float(const std::string& key, float fallback) const
{
    const Value* v = find(key);
    return v && v->type == Number ? (float)v->number : fallback;
}
bool Value::get_bool(const std::string& key, bool fallback) const
{
    const Value* v = find(key);
    if (v && v->type == Bool) return v->boolean;
    if (v && v->type == Number) return v->number != 0.0;
    return fallback;
}
std::string Value::get_string(const std::string& key, const std::string& fallback) const
{
    const Value* v = find(key);
    return v && v->type == String ? v->string : fallback;
}
bool Value::get_vector(const std::string& key, float* dst, size_t n) const
{
    const Value* v = find(key);
    if (!v || v->type != Array || v->array.size() != n) return false;
    for (const Value& e : v->array)
        if (e.type != Number) return false;
    for (size_t i = 0; i < n; i++) dst[i] = (float)v->array[i].number;
    return true;
}
Value parse(const std::string& text)
{
    Parser p(text);
    Value  v = p.value(0);
    p.skip();
    if (p.p != text.size()) p.fail("trailing characters");
    return v;
}
Value parse_file(const std::string& path)
{
    std::ifstream f(path, std::ios::binary);
    if (!f.is_open()) throw std::runtime_error("json: cannot open " + path);
    std::stringstream ss;
    ss << f.rdbuf();
    return parse(ss.str());
}
} // namespace json
} // namespace helios
