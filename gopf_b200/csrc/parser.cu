// Host-side restatement of the reference's equation-string logic
// (/root/reference/pf/rhsBuilder.go, /root/reference/pf/util.go).
#include "parser.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <deque>
#include <regex>
#include <stdexcept>

#include "host_util.h"

namespace gopf {
namespace parser {

static std::regex compile(const std::string& pattern) {
    try {
        return std::regex(pattern, std::regex::ECMAScript);
    } catch (const std::regex_error& e) {
        throw Error("parser: cannot compile regular expression '" + pattern + "': " + e.what());
    }
}

std::vector<std::string> go_find_all(const std::string& pattern, const std::string& s) {
    const std::regex rx = compile(pattern);
    std::vector<std::string> out;
    size_t pos = 0;
    long prev_end = -1;
    const size_t end = s.size();
    while (pos <= end) {
        std::smatch m;
        auto flags = std::regex_constants::match_default;
        if (pos > 0) flags |= std::regex_constants::match_prev_avail;
        if (!std::regex_search(s.begin() + pos, s.end(), m, rx, flags)) break;
        const size_t mstart = pos + (size_t)m.position(0);
        const size_t mend = mstart + (size_t)m.length(0);
        bool accept = true;
        if (mend == pos) {  // empty match at the search position
            if ((long)mstart == prev_end) accept = false;
            pos = pos + 1;
        } else {
            pos = mend;
        }
        prev_end = (long)mend;
        if (accept) out.push_back(m.str(0));
    }
    return out;
}

std::string go_find_string(const std::string& pattern, const std::string& s) {
    const std::regex rx = compile(pattern);
    std::smatch m;
    if (std::regex_search(s, m, rx)) return m.str(0);
    return "";
}

static bool find_power(const std::string& pattern, std::string* num) {
    static const std::regex rx("\\^(-?\\d+\\.?\\d*)");
    std::smatch m;
    if (!std::regex_search(pattern, m, rx)) return false;
    *num = m.str(1);
    return true;
}

double get_power(const std::string& pattern) {
    std::string num;
    if (!find_power(pattern, &num)) return 1.0;
    return std::strtod(num.c_str(), nullptr);
}

std::vector<std::string> split(const std::string& s, const std::string& delim) {
    std::vector<std::string> out;
    size_t start = 0;
    while (true) {
        size_t p = s.find(delim, start);
        if (p == std::string::npos) {
            out.push_back(s.substr(start));
            break;
        }
        out.push_back(s.substr(start, p - start));
        start = p + delim.size();
    }
    return out;
}

bool contains(const std::string& s, const std::string& sub) { return s.find(sub) != std::string::npos; }

std::string replace_all(std::string s, const std::string& from, const std::string& to) {
    if (from.empty()) return s;
    size_t p = 0;
    while ((p = s.find(from, p)) != std::string::npos) {
        s.replace(p, from.size(), to);
        p += to.size();
    }
    return s;
}

std::string strip_spaces(const std::string& s) { return replace_all(s, " ", ""); }

std::string sort_factors(const std::string& expr) {
    std::vector<std::string> parts = split(expr, "*");
    std::sort(parts.begin(), parts.end());  // bytewise, like Go sort.Strings
    std::string out;
    for (size_t i = 0; i < parts.size(); ++i) {
        if (i) out += "*";
        out += parts[i];
    }
    return out;
}

std::string get_field_name(const std::string& term, const std::vector<std::string>& names) {
    std::string field;
    for (const std::string& f : names) {
        if (contains(term, f)) {
            const std::string without = replace_all(term, f, "");
            bool ok = true;
            for (const std::string& f1 : names) {
                if (contains(without, f1)) {
                    ok = false;
                    break;
                }
            }
            if (ok && f.size() > field.size()) field = f;
        }
    }
    return field;
}

static std::string first_delimiter(const std::string& value, const std::vector<std::string>& delims) {
    for (const std::string& d : delims)
        if (!value.empty() && value.substr(0, 1) == d) return d;
    return "";
}

std::vector<SubStringDelimiter> split_on_many(const std::string& value, const std::vector<std::string>& delims) {
    std::vector<SubStringDelimiter> out;
    std::deque<SubStringDelimiter> queue;
    queue.push_back({value, first_delimiter(value, delims)});
    std::string all;
    for (const std::string& d : delims) all += d;
    while (!queue.empty()) {
        SubStringDelimiter cur = queue.front();
        queue.pop_front();
        if (cur.SubString.find_first_of(all) == std::string::npos) {
            out.push_back(cur);
            continue;
        }
        std::string delim = delims[0];
        for (const std::string& d : delims) {
            if (contains(cur.SubString, d)) {
                delim = d;
                break;
            }
        }
        std::vector<std::string> splits;
        for (const std::string& s : split(cur.SubString, delim))
            if (!s.empty()) splits.push_back(s);
        if (splits.empty()) continue;  // the reference would index splits[0] and panic
        queue.push_back({splits[0], cur.PreceedingDelimiter});
        for (size_t i = 1; i < splits.size(); ++i) queue.push_back({splits[i], delim});
    }
    return out;
}

bool is_bilinear(const std::string& term, const std::string& field, const std::vector<std::string>& names) {
    if (go_find_all(field, term).size() != 1) return false;
    for (const std::string& f : names) {
        if (f == field) continue;
        if (!go_find_all(f, term).empty()) return false;
    }
    const std::string res = go_find_string(field + "*[^/\\*]*", term);
    std::string num;
    if (!find_power(res, &num)) return true;
    char* endp = nullptr;
    const double power = std::strtod(num.c_str(), &endp);
    if (endp == num.c_str()) return true;
    return std::fabs(power - 1.0) < 1e-10;
}

std::string get_non_linear_field_expressions(const std::string& pattern, const std::string& field,
                                             const std::vector<std::string>& names) {
    std::string expr;
    for (const std::string& fn : names) {
        if (fn == field && is_bilinear(pattern, field, names)) continue;
        const std::string res = go_find_string(fn + "[^\\*]*", pattern);
        if (!res.empty()) expr += res + "*";
    }
    if (expr.size() > 1) return expr.substr(0, expr.size() - 1);
    return expr;
}

std::string field_name_from_leibniz(const std::string& leibniz) {
    if (leibniz.size() <= 3) throw Error("rhsbuilder: Length of the Leibniz formatted string has to be at least 3");
    if (leibniz.substr(0, 1) != "d" || leibniz.substr(leibniz.size() - 3) != "/dt")
        throw Error("rhsbuilder: Passed string is not a leibniz formatted string");
    return leibniz.substr(1, leibniz.size() - 4);
}

std::vector<std::string> known_prefixes() { return {"-", "LAP^4", "LAP^2", "LAP", "*"}; }

static bool has_prefix(const std::string& s, const std::string& p) { return s.compare(0, p.size(), p) == 0; }

std::vector<std::string> get_known_prefixes(std::string s) {
    std::vector<std::string> pref;
    std::vector<std::string> prefixes = known_prefixes();
    while (!prefixes.empty()) {
        const std::string p = prefixes.front();
        prefixes.erase(prefixes.begin());
        if (has_prefix(s, p)) {
            pref.push_back(p);
            prefixes = known_prefixes();
            s = s.substr(p.size());
        }
    }
    return pref;
}

std::string remove_known_prefixes(std::string s) {
    std::vector<std::string> prefixes = known_prefixes();
    while (!prefixes.empty()) {
        const std::string p = prefixes.front();
        prefixes.erase(prefixes.begin());
        if (has_prefix(s, p)) {
            s = s.substr(p.size());
            prefixes = known_prefixes();
        }
    }
    return s;
}

}  // namespace parser
}  // namespace gopf
