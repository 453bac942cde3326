// barcodes.cpp -- host-side barcode-table loader: the C++ twin of the
// reference's preprocess_bc_file (src/fileio.jl:7-72).  FASTA by extension (.fasta / .fa), otherwise a
// delimited table (',' for .csv, tab for anything else) with the columns Full_seq, ID and
// Full_annotation; bases whose annotation character is not 'B' are dropped, the rest is upper-cased,
// U -> T, optionally complemented (only ATGCN are mapped) and reversed.  The output is exactly what
// bdx_barcode_set wants: concatenated bytes, offsets, lengths_no_N -- plus the IDs for file naming.
#include <cstdio>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "../../include/bdx.h"

struct bdx_barcode_table {
    std::vector<std::string> seqs, ids;
    std::vector<uint8_t> bytes;
    std::vector<int32_t> offsets, lengths_no_n;
};

namespace {

thread_local std::string g_bc_err;

bool ends_with_ci(const std::string &s, const char *suffix)
{
    const size_t n = strlen(suffix);
    if (s.size() < n) return false;
    for (size_t i = 0; i < n; i++) {
        char c = s[s.size() - n + i];
        if (c >= 'A' && c <= 'Z') c = (char)(c - 'A' + 'a');
        if (c != suffix[i]) return false;
    }
    return true;
}

bool read_file(const char *path, std::string &out)
{
    FILE *f = fopen(path, "rb");
    if (!f) return false;
    char buf[1 << 16];
    size_t n;
    while ((n = fread(buf, 1, sizeof(buf), f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}

// the characters (UTF-8 code points) of a string: the reference zips, reverses and counts characters
std::vector<std::string> chars_of(const std::string &s)
{
    std::vector<std::string> out;
    for (size_t i = 0; i < s.size();) {
        const unsigned char c = (unsigned char)s[i];
        size_t n = c < 0x80 ? 1 : (c >> 5) == 6 ? 2 : (c >> 4) == 14 ? 3 : (c >> 3) == 30 ? 4 : 1;
        if (i + n > s.size()) n = s.size() - i;
        out.emplace_back(s, i, n);
        i += n;
    }
    return out;
}

bool is_space(char c) { return c == ' ' || c == '\t' || c == '\n' || c == '\v' || c == '\f' || c == '\r'; }

std::string strip(const std::string &s)
{
    size_t a = 0, b = s.size();
    while (a < b && is_space(s[a])) a++;
    while (b > a && is_space(s[b - 1])) b--;
    return s.substr(a, b - a);
}

// eachline: split at "\n", strip one trailing "\r"
std::vector<std::string> lines_of(const std::string &text)
{
    std::vector<std::string> out;
    size_t pos = 0;
    while (pos < text.size()) {
        size_t e = text.find('\n', pos);
        size_t next;
        if (e == std::string::npos) {
            e = text.size();
            next = e;
        } else {
            next = e + 1;
            if (e > pos && text[e - 1] == '\r') e--;
        }
        out.emplace_back(text, pos, e - pos);
        pos = next;
    }
    return out;
}

// delimited table: quoted fields ("" escapes a quote, may hold delimiters and newlines), rows end at
// "\n" or "\r\n", empty rows are skipped (CSV.jl defaults)
std::vector<std::vector<std::string>> parse_table(const std::string &text, char delim)
{
    std::vector<std::vector<std::string>> rows;
    std::vector<std::string> row;
    std::string field;
    bool in_quotes = false, field_started = false, row_has_data = false;
    size_t i = 0;
    if (text.size() >= 3 && (unsigned char)text[0] == 0xEF && (unsigned char)text[1] == 0xBB && (unsigned char)text[2] == 0xBF)
        i = 3;   // byte-order mark
    auto end_field = [&]() {
        row.push_back(field);
        field.clear();
        field_started = false;
    };
    auto end_row = [&]() {
        end_field();
        if (row_has_data) rows.push_back(row);
        row.clear();
        row_has_data = false;
    };
    for (; i < text.size(); i++) {
        const char c = text[i];
        if (in_quotes) {
            if (c == '"') {
                if (i + 1 < text.size() && text[i + 1] == '"') {
                    field.push_back('"');
                    i++;
                } else {
                    in_quotes = false;
                }
            } else {
                field.push_back(c);
            }
            continue;
        }
        if (c == '"' && !field_started) {
            in_quotes = true;
            field_started = true;
            row_has_data = true;
        } else if (c == delim) {
            row_has_data = true;
            end_field();
        } else if (c == '\n') {
            end_row();
        } else if (c == '\r' && i + 1 < text.size() && text[i + 1] == '\n') {
            // part of "\r\n"
        } else {
            field.push_back(c);
            field_started = true;
            row_has_data = true;
        }
    }
    if (row_has_data || !field.empty()) end_row();
    return rows;
}

char complement_of(char c)
{
    switch (c) {   // fileio.jl:58-63
    case 'A': return 'T'; case 'T': return 'A'; case 'G': return 'C'; case 'C': return 'G';
    case 'a': return 't'; case 't': return 'a'; case 'g': return 'c'; case 'c': return 'g';
    default: return c;   // N, n and everything else map to themselves
    }
}

int fail(int code, const std::string &msg)
{
    g_bc_err = msg;
    return code;
}

}  // namespace

extern "C" const char *bdx_barcode_table_error(void) { return g_bc_err.c_str(); }

extern "C" int bdx_barcode_table_load(const char *path, int complement, int rev, bdx_barcode_table **out)
{
    if (!path || !out) return fail(BDX_ERR_INVALID, "null argument");
    *out = nullptr;
    std::string text;
    if (!read_file(path, text)) return fail(BDX_ERR_INVALID, std::string("cannot open barcode file ") + path);
    bdx_barcode_table *t = new (std::nothrow) bdx_barcode_table();
    if (!t) return fail(BDX_ERR_NOMEM, "out of memory");
    std::vector<std::string> annotations;
    const std::string p(path);
    if (ends_with_ci(p, ".fasta") || ends_with_ci(p, ".fa")) {            // fileio.jl:11-32
        std::string current;
        for (const std::string &line : lines_of(text)) {
            if (!line.empty() && line[0] == '>') {
                if (!current.empty()) {
                    t->seqs.push_back(current);
                    current.clear();
                }
                std::string id = strip(line.substr(1));
                for (size_t k = 0; k < id.size(); k++)
                    if (is_space(id[k])) {                                  // replace(.., r"\s.*$" => "")
                        id.resize(k);
                        break;
                    }
                t->ids.push_back(id);
            } else {
                current += strip(line);
            }
        }
        if (!current.empty()) t->seqs.push_back(current);
        for (const std::string &s : t->seqs) annotations.emplace_back(chars_of(s).size(), 'B');
    } else {                                                                // fileio.jl:33-40
        const char delim = ends_with_ci(p, ".csv") ? ',' : '\t';
        const auto rows = parse_table(text, delim);
        int i_seq = -1, i_id = -1, i_ann = -1;
        if (!rows.empty())
            for (size_t k = 0; k < rows[0].size(); k++) {
                if (rows[0][k] == "Full_seq") i_seq = (int)k;
                if (rows[0][k] == "ID") i_id = (int)k;
                if (rows[0][k] == "Full_annotation") i_ann = (int)k;
            }
        if (i_seq < 0 || i_id < 0 || i_ann < 0) {
            delete t;
            return fail(BDX_ERR_INVALID, std::string("barcode table ") + path + " needs columns Full_seq, ID, Full_annotation");
        }
        for (size_t r = 1; r < rows.size(); r++) {
            auto cell = [&](int k) { return k < (int)rows[r].size() ? rows[r][k] : std::string(); };
            t->seqs.push_back(cell(i_seq));
            t->ids.push_back(cell(i_id));
            annotations.push_back(cell(i_ann));
        }
    }

    for (size_t i = 0; i < t->seqs.size(); i++) {
        const auto sc = chars_of(t->seqs[i]), ac = chars_of(annotations[i]);
        if (sc.size() != ac.size()) {                                       // fileio.jl:45-47
            const std::string id = i < t->ids.size() ? t->ids[i] : std::string();
            delete t;
            return fail(BDX_ERR_INVALID, "Length mismatch between sequence and annotation for ID: " + id);
        }
        std::vector<std::string> kept;
        for (size_t k = 0; k < sc.size(); k++)
            if (ac[k] == "B") kept.push_back(sc[k]);                        // :49
        for (std::string &c : kept) {
            if (c.size() == 1) {
                char ch = c[0];
                if (ch >= 'a' && ch <= 'z') ch = (char)(ch - 'a' + 'A');    // uppercase (ASCII; other scripts kept as is)
                if (ch == 'U') ch = 'T';                                    // :55
                if (complement) ch = complement_of(ch);                     // :57-64
                c[0] = ch;
            }
        }
        std::string s;
        if (rev)                                                            // :65-67
            for (size_t k = kept.size(); k-- > 0;) s += kept[k];
        else
            for (const std::string &c : kept) s += c;
        int no_n = 0;
        for (const std::string &c : kept) no_n += c != "N";                 // :69
        t->seqs[i] = s;
        t->lengths_no_n.push_back(no_n);
    }
    t->offsets.push_back(0);
    for (const std::string &s : t->seqs) {
        t->bytes.insert(t->bytes.end(), s.begin(), s.end());
        if (t->bytes.size() > 0x7FFFFFF0u) {
            delete t;
            return fail(BDX_ERR_TOO_LARGE, "barcode table too large");
        }
        t->offsets.push_back((int32_t)t->bytes.size());
    }
    if (t->bytes.empty()) t->bytes.push_back(0);   // keep data() non-null
    *out = t;
    return BDX_OK;
}

extern "C" void bdx_barcode_table_destroy(bdx_barcode_table *t) { delete t; }
extern "C" int32_t bdx_barcode_table_count(const bdx_barcode_table *t) { return t ? (int32_t)t->seqs.size() : 0; }
extern "C" int32_t bdx_barcode_table_id_count(const bdx_barcode_table *t) { return t ? (int32_t)t->ids.size() : 0; }
extern "C" const uint8_t *bdx_barcode_table_bytes(const bdx_barcode_table *t) { return t ? t->bytes.data() : nullptr; }
extern "C" const int32_t *bdx_barcode_table_offsets(const bdx_barcode_table *t) { return t ? t->offsets.data() : nullptr; }
extern "C" const int32_t *bdx_barcode_table_lengths_no_n(const bdx_barcode_table *t)
{
    return t ? t->lengths_no_n.data() : nullptr;
}
extern "C" const char *bdx_barcode_table_id(const bdx_barcode_table *t, int32_t i)
{
    return (t && i >= 0 && i < (int32_t)t->ids.size()) ? t->ids[(size_t)i].c_str() : nullptr;
}
