// qcir_io.cpp -- gate-name table and QASM reader of the minimal QCir (qcir.hpp).
// Behaviour follows the reference's src/qcir/gate_type.cpp:16-70 and
// src/qcir/qcir_reader.cpp:47-120 (quirks included: exactly six header tokens, `//` comments,
// unknown gates are skipped silently, a leading run of 'c' counts controls).
#include <cctype>
#include <fstream>
#include <sstream>
#include <string_view>

#include "./qcir.hpp"

namespace qsyn::qcir {

namespace {

std::string lower(std::string s) {
    for (auto& ch : s) ch = static_cast<char>(std::tolower(static_cast<unsigned char>(ch)));
    return s;
}

std::optional<Operation> basic_operation(std::string const& name, std::vector<dvlab::Phase> const& params) {
    using dvlab::Phase;
    if (params.empty()) {
        if (name == "id") return IdGate();
        if (name == "h") return HGate();
        if (name == "swap") return SwapGate();
        if (name == "ecr") return ECRGate();
        struct Fixed {
            char const* name;
            char axis;
            int num, den;
        };
        static Fixed const table[] = {
            {"z", 'z', 1, 1},     {"s", 'z', 1, 2},     {"sdg", 'z', -1, 2},  {"t", 'z', 1, 4},   {"tdg", 'z', -1, 4},
            {"x", 'x', 1, 1},     {"not", 'x', 1, 1},   {"sx", 'x', 1, 2},    {"x_1_2", 'x', 1, 2}, {"sxdg", 'x', -1, 2},
            {"tx", 'x', 1, 4},    {"txdg", 'x', -1, 4}, {"y", 'y', 1, 1},     {"sy", 'y', 1, 2},  {"y_1_2", 'y', 1, 2},
            {"sydg", 'y', -1, 2}, {"ty", 'y', 1, 4},    {"tydg", 'y', -1, 4},
        };
        for (auto const& f : table) {
            if (name != f.name) continue;
            Phase const ph(f.num, f.den);
            if (f.axis == 'z') return PZGate(ph);
            if (f.axis == 'x') return PXGate(ph);
            return PYGate(ph);
        }
    }
    if (params.size() == 1) {
        if (name == "p" || name == "pz") return PZGate(params[0]);
        if (name == "px") return PXGate(params[0]);
        if (name == "py") return PYGate(params[0]);
        if (name == "rz") return RZGate(params[0]);
        if (name == "rx") return RXGate(params[0]);
        if (name == "ry") return RYGate(params[0]);
    }
    return std::nullopt;
}

// token starting at the first non-delimiter at/after pos; returns the position of the delimiter
// that ends it (npos if it runs to the end, or if there is no token).  Views into `s`: the reader allocates nothing per token.
constexpr std::string_view kBlank = " \t\n\v\f\r";
size_t next_token(std::string_view s, std::string_view& tok, size_t pos = 0, std::string_view delim = kBlank) {
    size_t const begin = pos == std::string_view::npos ? pos : s.find_first_not_of(delim, pos);
    if (begin == std::string_view::npos) {
        tok = {};
        return begin;
    }
    size_t const end = s.find_first_of(delim, begin);
    tok              = s.substr(begin, end == std::string_view::npos ? end : end - begin);
    return end;
}

std::string_view trim(std::string_view s) {
    size_t const b = s.find_first_not_of(kBlank);
    if (b == std::string_view::npos) return {};
    return s.substr(b, s.find_last_not_of(kBlank) + 1 - b);
}

std::optional<Operation> build_operation(std::string str, std::vector<dvlab::Phase> const& params);

}  // namespace

std::optional<Operation> str_to_operation(std::string str, std::vector<dvlab::Phase> const& params) {
    // A circuit file names a handful of distinct parameter-free gates thousands of times: the (immutable, shared) Operation
    // of a name is built once per thread.  Parametrised gates are built per call.
    if (params.empty()) {
        static thread_local std::unordered_map<std::string, std::optional<Operation>> memo;
        auto it = memo.find(str);
        if (it == memo.end()) {
            if (memo.size() >= 1024) memo.clear();
            it = memo.emplace(str, build_operation(str, params)).first;
        }
        return it->second;
    }
    return build_operation(std::move(str), params);
}

namespace {
std::optional<Operation> build_operation(std::string str, std::vector<dvlab::Phase> const& params) {
    str                  = lower(std::move(str));
    size_t const n_ctrls = str.find_first_not_of('c');
    if (n_ctrls == std::string::npos) return std::nullopt;  // the reference throws here ("c", "cc", ...)
    auto basic = basic_operation(str.substr(n_ctrls), params);
    if (!basic.has_value()) return std::nullopt;
    if (n_ctrls > 0) return ControlGate(*basic, n_ctrls);
    return basic;
}
}  // namespace

std::optional<QCir> from_qasm(std::filesystem::path const& filepath) {
    std::ifstream file{filepath};
    if (!file.is_open()) {
        spdlog::error("Cannot open the QASM file \"{}\"!!", filepath.string());
        return std::nullopt;
    }
    auto qc = from_qasm_stream(file, filepath.string());
    if (qc.has_value()) qc->set_filename(filepath.string());
    return qc;
}

std::optional<QCir> from_qasm_stream(std::istream& file, std::string const& /*name_for_messages*/) {
    // OPENQASM 2.0; include "qelib1.inc"; qreg q[N];  -- six whitespace separated tokens
    std::string word;
    for (int i = 0; i < 6; ++i) file >> word;
    size_t const lb     = word.find('[');
    size_t const nqubit = std::stoul(word.substr(lb + 1, word.size() - lb - 3));
    QCir qcir{nqubit};

    std::string raw;
    QubitIdList qubit_ids;
    std::getline(file, raw);  // rest of the qreg line
    while (std::getline(file, raw)) {
        std::string_view const line = trim(std::string_view(raw).substr(0, raw.find("//")));
        if (line.empty()) continue;
        std::string_view type, phase_str = "0", scratch;
        size_t const type_end = next_token(line, type);
        if (next_token(line, scratch, 0, "(") != std::string_view::npos) {
            size_t const stop = next_token(line, type, 0, "(");
            next_token(line, phase_str, stop + 1, ")");
        }
        if (type == "creg" || type == "qreg" || type.empty()) continue;

        qubit_ids.clear();
        std::string_view token, id_str;
        size_t pos = next_token(line, token, type_end, ",");
        while (!token.empty()) {
            next_token(token, id_str, next_token(token, id_str, 0, "[") + 1, "]");
            bool const numeric = !id_str.empty() && std::all_of(id_str.begin(), id_str.end(), [](unsigned char c) { return std::isdigit(c); });
            size_t id = nqubit;
            if (numeric && id_str.size() <= 18) {
                id = 0;
                for (char ch : id_str) id = 10 * id + static_cast<size_t>(ch - '0');
            }
            if (id >= nqubit) {
                spdlog::error("invalid qubit id on line {}!!", line);
                return std::nullopt;
            }
            qubit_ids.emplace_back(id);
            pos = next_token(line, token, pos, ",");
        }
        if (!QCirGate::qubit_id_is_unique(qubit_ids)) {
            spdlog::error("duplicate qubit id on line {}!!", line);
            return std::nullopt;
        }
        if (auto op = str_to_operation(std::string(type)); op.has_value() && op->get_num_qubits() == qubit_ids.size()) {
            qcir.append(*op, qubit_ids);
            continue;
        }
        auto const phase = dvlab::Phase::from_string(std::string(phase_str));
        if (!phase.has_value()) {
            spdlog::error("invalid phase on line {}!!", line);
            return std::nullopt;
        }
        if (auto op = str_to_operation(std::string(type), {*phase}); op.has_value() && op->get_num_qubits() == qubit_ids.size()) {
            qcir.append(*op, qubit_ids);
            continue;
        }
        // anything else is skipped silently, like the reference
    }
    return qcir;
}

std::optional<QCir> from_file(std::filesystem::path const& filepath) {
    auto const ext = filepath.extension().string();
    if (ext == ".qasm") return from_qasm(filepath);
    spdlog::error("File format \"{}\" is not supported!!", ext);
    return std::nullopt;
}

}  // namespace qsyn::qcir
