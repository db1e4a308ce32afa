// Equation-string parser: host-side restatement of the reference's
// pf/rhsBuilder.go + pf/util.go string logic.  Names follow the Go functions.
#pragma once
#include <string>
#include <utility>
#include <vector>

namespace gopf {
namespace parser {

// Go regexp.FindAllString(s, -1): empty matches abutting the preceding match are dropped.
std::vector<std::string> go_find_all(const std::string& pattern, const std::string& s);
std::string go_find_string(const std::string& pattern, const std::string& s);

double get_power(const std::string& pattern);                                        // util.go:67-80
std::string sort_factors(const std::string& expr);                                    // util.go:296-300
std::string get_field_name(const std::string& term, const std::vector<std::string>& names);  // util.go:82-104

struct SubStringDelimiter {  // util.go:134-139
    std::string SubString;
    std::string PreceedingDelimiter;
};
std::vector<SubStringDelimiter> split_on_many(const std::string& value, const std::vector<std::string>& delims);  // util.go:152-199

bool is_bilinear(const std::string& term, const std::string& field, const std::vector<std::string>& names);  // rhsBuilder.go:69-106
std::string get_non_linear_field_expressions(const std::string& pattern, const std::string& field,
                                             const std::vector<std::string>& names);  // util.go:17-36
std::string field_name_from_leibniz(const std::string& leibniz);                      // rhsBuilder.go:58-66
std::vector<std::string> known_prefixes();                                            // rhsBuilder.go:244-252
std::vector<std::string> get_known_prefixes(std::string s);                           // rhsBuilder.go:259-272
std::string remove_known_prefixes(std::string s);                                     // rhsBuilder.go:254-257,274-287
std::string strip_spaces(const std::string& s);
std::string replace_all(std::string s, const std::string& from, const std::string& to);
std::vector<std::string> split(const std::string& s, const std::string& delim);       // Go strings.Split
bool contains(const std::string& s, const std::string& sub);

}  // namespace parser
}  // namespace gopf
