// utility/json.h — minimal JSON document model + parser for the AssetCore material / scene descriptions
// (the reference reads them with the vendored nlohmann::json, external/AssetCore/src/loader/loader.cpp:166-568;
// only what those files use is supported: objects, arrays, strings with the standard escapes, numbers, booleans,
// null).  Parse errors throw std::runtime_error with the byte offset.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace helios
{
namespace json
{
class Value
{
public:
    enum Type
    {
        Null,
        Bool,
        Number,
        String,
        Array,
        Object
    };
    Type                                        type = Null;
    bool                                        boolean = false;
    double                                      number  = 0.0;
    std::string                                 string;
    std::vector<Value>                          array;
    std::vector<std::pair<std::string, Value>>  object; // insertion order kept; find() returns the LAST duplicate as nlohmann does

    const Value* find(const std::string& key) const;
    bool         is_number() const { return type == Number; }
    bool         is_string() const { return type == String; }
    // typed reads with a default used when the key is absent or of the wrong type
    float       get_float(const std::string& key, float fallback) const;
    bool        get_bool(const std::string& key, bool fallback) const;
    std::string get_string(const std::string& key, const std::string& fallback) const;
    // reads key as an array of exactly N numbers into dst; leaves dst untouched otherwise (JSON_PARSE_VECTOR, loader.cpp:12-27)
    bool get_vector(const std::string& key, float* dst, size_t n) const;
};
Value parse(const std::string& text);
Value parse_file(const std::string& path); // throws if the file cannot be opened
} // namespace json
} // namespace helios
