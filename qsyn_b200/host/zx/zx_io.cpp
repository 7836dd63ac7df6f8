// zx_io.cpp -- the `.zx` reader (src/zx/zx_io.cpp:83-436 of the reference): one vertex per line,
//   <I|O><id> [(<row>,<column>)] [<S|H><neighbor id> ...] [qubit id]
//   <Z|X|H><id> [(<row>,<column>)] [<S|H><neighbor id> ...] [phase]
// `//` starts a comment.  Error texts follow the reference.
#include <cctype>
#include <charconv>

#include "./zxgraph.hpp"

namespace qsyn::zx {

namespace {

struct VertexInfo {
    char type = 'Z';
    QubitIdType qubit = 0;
    float row = 0.f, column = 0.f;
    std::vector<std::pair<char, size_t>> neighbors;
    dvlab::Phase phase;
};

size_t get_token(std::string const& s, std::string& tok, size_t pos = 0, std::string const& delim = " \t\n\v\f\r") {
    size_t const begin = pos == std::string::npos ? pos : s.find_first_not_of(delim, pos);
    if (begin == std::string::npos) {
        tok.clear();
        return begin;
    }
    size_t const end = s.find_first_of(delim, begin);
    tok              = s.substr(begin, end - begin);
    return end;
}

std::string trim(std::string const& s) {
    size_t const b = s.find_first_not_of(" \t\n\v\f\r");
    if (b == std::string::npos) return "";
    return s.substr(b, s.find_last_not_of(" \t\n\v\f\r") + 1 - b);
}

template <typename T>
std::optional<T> parse_number(std::string const& s) {
    T v{};
    auto const [p, ec] = std::from_chars(s.data(), s.data() + s.size(), v);
    if (ec == std::errc{} && p == s.data() + s.size()) return v;
    return std::nullopt;
}

class Parser {
public:
    std::optional<std::vector<std::pair<size_t, VertexInfo>>> parse(std::istream& f) {
        std::vector<std::pair<size_t, VertexInfo>> storage;
        std::unordered_set<size_t> ids;
        std::unordered_set<QubitIdType> taken_inputs;
        QubitIdType max_in = 0, max_out = 0;
        _line_no = 1;
        for (std::string line; std::getline(f, line); ++_line_no) {
            line = trim(line.substr(0, line.find("//")));
            if (line.empty()) continue;
            std::vector<std::string> tokens;
            if (!tokenize(line, tokens)) return std::nullopt;
            VertexInfo info;
            char const type = static_cast<char>(std::toupper(static_cast<unsigned char>(tokens[0][0])));
            if (type == 'G') return fail("ground vertices are not supported yet!!");
            if (std::string("IOZXH").find(type) == std::string::npos) return fail(std::format("unsupported vertex type ({})!!", type));
            std::string const id_string = tokens[0].substr(1);
            if (id_string.empty()) return fail(std::format("Missing vertex ID after vertex type declaration ({})!!", type));
            auto const id = parse_number<size_t>(id_string);
            if (!id) return fail(std::format("vertex ID ({}) is not an unsigned integer!!", id_string));
            if (ids.contains(*id)) return fail(std::format("duplicated vertex ID ({})!!", *id));
            info.type = type;
            if (type == 'H') info.phase = dvlab::Phase(1);
            if (type == 'I' || type == 'O') {
                QubitIdType& next = type == 'I' ? max_in : max_out;
                if (auto const q = parse_number<int>(tokens.back()); q.has_value() && tokens.size() > 3) {
                    tokens.pop_back();
                    info.qubit = *q;
                    next       = std::max(next, info.qubit);
                } else {
                    info.qubit = next++;
                }
                if (type == 'I') {
                    if (taken_inputs.contains(info.qubit)) return fail(std::format("duplicated input qubit ID ({})!!", info.qubit));
                    taken_inputs.insert(info.qubit);
                }
                info.row = static_cast<float>(info.qubit);
            } else {
                if (tokens.size() > 3) {
                    if (auto const ph = dvlab::Phase::from_string(tokens.back()); ph.has_value()) {
                        tokens.pop_back();
                        info.phase = *ph;
                    }
                }
                info.row = 0;
            }
            if (tokens[1] != "-") {
                auto const r = parse_number<float>(tokens[1]);
                if (!r) return fail(std::format("row ({}) is not an floating-point number!!", tokens[1]));
                info.row = *r;
            }
            if (tokens[2] == "-") {
                info.column = 0;
            } else {
                auto const c = parse_number<float>(tokens[2]);
                if (!c) return fail(std::format("column ({}) is not an floating-point number!!", tokens[2]));
                info.column = *c;
            }
            for (size_t i = 3; i < tokens.size(); ++i) {
                char const et = static_cast<char>(std::toupper(static_cast<unsigned char>(tokens[i][0])));
                if (et != 'S' && et != 'H') return fail(std::format("unsupported edge type ({})!!", et));
                std::string const nb = tokens[i].substr(1);
                if (nb.empty()) return fail(std::format("Missing neighbor vertex ID after edge type declaration ({})!!", et));
                auto const nid = parse_number<unsigned>(nb);
                if (!nid) return fail(std::format("neighbor vertex ID ({}) is not an unsigned integer!!", nb));
                info.neighbors.emplace_back(et, *nid);
            }
            ids.insert(*id);
            storage.emplace_back(*id, std::move(info));
        }
        return storage;
    }

private:
    size_t _line_no = 1;

    std::nullopt_t fail(std::string const& what) {
        spdlog::error("Error: failed to read line {}!!", _line_no);
        spdlog::error("{}", what);
        return std::nullopt;
    }
    bool tokenize(std::string const& line, std::vector<std::string>& tokens) {
        std::string token;
        size_t pos = get_token(line, token);
        tokens.push_back(token);
        size_t const lp = line.find_first_of('(', pos);
        size_t const rp = line.find_first_of(')', lp == std::string::npos ? 0 : lp);
        if (lp == std::string::npos && rp == std::string::npos) {
            tokens.emplace_back("-");
            tokens.emplace_back("-");
        } else if (rp == std::string::npos) {
            fail("missing closing parenthesis!!");
            return false;
        } else if (lp == std::string::npos) {
            fail("missing opening parenthesis!!");
            return false;
        } else {
            pos = get_token(line, token, lp + 1, ",");
            if (pos == std::string::npos) {
                fail("missing comma between declaration of qubit and column!!");
                return false;
            }
            token = trim(token);
            if (token.empty()) {
                fail("missing argument before comma!!");
                return false;
            }
            tokens.push_back(token);
            get_token(line, token, pos + 1, ")");
            token = trim(token);
            if (token.empty()) {
                fail("missing argument before right parenthesis!!");
                return false;
            }
            tokens.push_back(token);
            pos = rp + 1;
        }
        pos = get_token(line, token, pos);
        while (!token.empty()) {
            tokens.push_back(token);
            pos = get_token(line, token, pos);
        }
        return true;
    }
};

}  // namespace

std::optional<ZXGraph> from_zx(std::istream& in, bool keep_id) {
    Parser parser;
    auto const storage = parser.parse(in);
    if (!storage) return std::nullopt;
    ZXGraph graph;
    std::unordered_map<size_t, ZXVertex*> id2v;
    for (auto const& [id, info] : *storage) {
        ZXVertex* v = nullptr;
        switch (info.type) {
            case 'I': v = graph.add_input(info.qubit, info.row, info.column); break;
            case 'O': v = graph.add_output(info.qubit, info.row, info.column); break;
            case 'Z': v = graph.add_vertex(VertexType::z, info.phase, info.row, info.column); break;
            case 'X': v = graph.add_vertex(VertexType::x, info.phase, info.row, info.column); break;
            default: v = graph.add_vertex(VertexType::h_box, info.phase, info.row, info.column); break;
        }
        if (v == nullptr) return std::nullopt;
        if (keep_id) v->set_id(id);
        id2v[id] = v;
    }
    for (auto const& [vid, info] : *storage) {
        for (auto const& [type, nbid] : info.neighbors) {
            if (!id2v.contains(nbid)) {
                spdlog::error("failed to build the graph: cannot find vertex with ID {}!!", nbid);
                return std::nullopt;
            }
            auto const etype = type == 'S' ? EdgeType::simple : EdgeType::hadamard;
            if (graph.is_neighbor(id2v[vid], id2v[nbid], etype)) continue;
            graph.add_edge(id2v[vid], id2v[nbid], etype);
        }
    }
    return graph;
}

std::optional<ZXGraph> from_zx(std::filesystem::path const& path, bool keep_id) {
    std::ifstream in(path);
    if (!in.is_open()) {
        spdlog::error("Cannot open the file \"{}\"!!", path.string());
        return std::nullopt;
    }
    return from_zx(in, keep_id);
}

}  // namespace qsyn::zx
