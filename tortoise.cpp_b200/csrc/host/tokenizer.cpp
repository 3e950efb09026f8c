// tokenizer.cpp -- the reference's tokenizer, restated: vocabulary scraped from the
// HuggingFace tokenizer.json by the same "poor-man's JSON" scan (common.cpp:166-262), GPT-2
// style regex word split (common.cpp:268-280) and greedy longest-match lookup
// (common.cpp:320-336).  Bit-exact ids are part of the drop-in contract (SURVEY A-14):
// the scan's quirks -- a bare (unquoted) value runs to the next ',' or '}', which swallows
// "[STOP]":0 into the value of "vocab"; non-integer values are dropped because stoi throws BEFORE the map slot is created (C++17 ordering,
// which is what the reference's default build uses) -- are reproduced, not fixed.
#include <cstdio>
#include <fstream>
#include <iterator>
#include <map>
#include <regex>
#include <string>
#include <vector>

namespace tts_host {

static std::string replace_all(std::string s, const std::string &from, const std::string &to) {
  size_t pos = 0;
  while ((pos = s.find(from, pos)) != std::string::npos) {
    s.replace(pos, from.length(), to);
    pos += to.length();
  }
  return s;
}

bool scrape_vocab(const std::string &path, std::map<std::string, int32_t> &vocab) {
  vocab.clear();
  std::ifstream ifs(path);
  if (!ifs) return false;
  const std::string js((std::istreambuf_iterator<char>(ifs)), std::istreambuf_iterator<char>());
  if (js.empty() || js[0] != '{') return true;
  const int n = int(js.size());
  auto at = [&](int i) -> char { return i < n ? js[i] : '\0'; };
  std::string key, val;
  bool have_key = false;   // a key string has been closed, value pending
  bool quoted = false;     // scanning inside "..."
  auto commit = [&]() {
    key = replace_all(key, "\\u0120", " ");
    key = replace_all(key, "\\u010a", "\n");
    key = replace_all(key, "\\\"", "\"");
    try {
      const int v = std::stoi(val);  // throws first: no slot is created for non-integers
      vocab[key] = v;
    } catch (...) {
    }
    key.clear();
    val.clear();
    quoted = false;
  };
  for (int i = 1; i < n; ++i) {
    const char ch = js[i];
    if (!quoted) {
      if (ch == ' ') continue;
      if (ch == '"') {
        quoted = true;
        continue;
      }
      continue;  // anything else outside quotes is ignored
    } else if (ch == '\\' && i + 1 < n) {
      (have_key ? val : key) += ch;  // keep the backslash, then the escaped character below
      ++i;
    } else if (ch == '"') {
      if (!have_key) {
        have_key = true;
        ++i;
        while (at(i) == ' ') ++i;
        ++i;  // the ':'
        while (at(i) == ' ') ++i;
        if (at(i) == '"') {  // string value follows: keep scanning it as a quoted token
          quoted = true;
          continue;
        }
        while (at(i) != ',' && at(i) != '}' && i < n) val += js[i++];
        have_key = false;
      } else {
        have_key = false;
      }
      commit();
      continue;
    }
    (have_key ? val : key) += js[i];
  }
  return true;
}

static void split_words(std::string str, std::vector<std::string> &words) {
  static const std::regex re(
      R"(\[SPACE\]|\[UNK\]|\[STOP\]|'s|'t|'re|'ve|'m|'ll|'d| ?[[:alpha:]]+| ?[[:digit:]]+| ?[^\s\[\][:alpha:][:digit:]]+|\s+(?!\S)|\s+)");
  std::smatch m;
  while (std::regex_search(str, m, re)) {
    words.push_back(m[0]);
    str = m.suffix();
  }
}

// message -> [255] + greedy ids + [0]   (main.cpp:6559-6567)
std::vector<int32_t> tokenize_message(const std::map<std::string, int32_t> &vocab, std::string message,
                                      bool warn) {
  message = replace_all(message, " ", "[SPACE]");
  std::vector<std::string> words;
  split_words(message, words);
  std::vector<int32_t> ids;
  ids.push_back(255);
  for (const std::string &w : words) {
    const int len = int(w.size());
    int i = 0;
    while (i < len) {
      bool hit = false;
      for (int j = len - 1; j >= i; --j) {
        auto it = vocab.find(w.substr(i, j - i + 1));
        if (it != vocab.end()) {
          ids.push_back(it->second);
          i = j + 1;
          hit = true;
          break;
        }
      }
      if (!hit) {
        if (warn) fprintf(stderr, "gpt_tokenize: unknown token '%s'\n", w.substr(i, 1).c_str());
        ++i;
      }
    }
  }
  ids.push_back(0);
  return ids;
}

}  // namespace tts_host
