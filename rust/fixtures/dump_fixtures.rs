// dump_fixtures.rs — writes reference-generated golden fixtures for searchlite-b200's oracle and CUDA engine.
//
// NOT compiled in the searchlite-b200 repository (no Rust toolchain there).  Copy this file to
// `searchlite-core/tests/dump_fixtures.rs` of davidkelley/searchlite and run
//
//   SLG_FIXTURE_DIR=/tmp/ref_fixtures cargo test --release --features vectors --test dump_fixtures -- --nocapture
//
// It builds four corpora through the reference's own writer, searches them through `IndexReader::search` (the call the
// GPU engine replaces, api/reader.rs:2539) under every ExecutionStrategy, and writes per corpus
//
//   ref_<name>.json      { k1, b, docs: [{id, body, lang?, year?, embedding?}], cases: [{query, terms, must, filter?, limit,
//                          execution, vector?, hits: [{doc_id, score, score_bits, vector_score?}], total_hits_estimate}] }
//   ref_index_<name>/    the index directory itself (MANIFEST.json, *.terms, *.post, *.fast, *.meta, *_vectors/*.bin):
//                        reference-WRITTEN segment files for slg_load_index_dir
//
// tests/test_reference_fixtures.py (searchlite-b200) consumes them.  Corpora:
//   pruning    tests/pruning.rs:45-104 — StdRng(42), 40 docs of 6 tokens over a 7-word vocabulary, k1 1.2 b 0.75, bmw block 4
//   query_ast  tests/query_ast.rs:52-58 — the literal 5-doc corpus with `lang` / `year` fast fields, k1 0.9 b 0.4
//   zipf       600 docs, 120-word vocabulary, Zipf-ish token draw from an explicit 64-bit LCG (reproducible without `rand`),
//              OR / AND / filtered queries, limits 5 and 40
//   hybrid     200 docs of `zipf` shape with 8-d unit vectors; BM25 candidates reranked with alpha 0.5 / 0.2 / 0.0
//              (HNSW ef_search >= #vectors, so the vector side is exhaustive — SURVEY.md §8c)
use std::collections::BTreeMap;
use std::fs;
use std::path::{Path, PathBuf};

use rand::rngs::StdRng;
use rand::{seq::SliceRandom, Rng, SeedableRng};
use searchlite_core::api::types::{
  Document, ExecutionStrategy, IndexOptions, KeywordField, NumericField, Query, QueryNode, Schema, SearchRequest, StorageType,
};
use searchlite_core::api::{Filter, Index};
use serde_json::{json, Value};

fn out_dir() -> PathBuf {
  let d = PathBuf::from(std::env::var("SLG_FIXTURE_DIR").unwrap_or_else(|_| "target/ref_fixtures".into()));
  fs::create_dir_all(&d).unwrap();
  d
}

fn opts(path: &Path, k1: f32, b: f32) -> IndexOptions {
  IndexOptions {
    path: path.to_path_buf(),
    create_if_missing: true,
    enable_positions: true,
    bm25_k1: k1,
    bm25_b: b,
    storage: StorageType::Filesystem,
    #[cfg(feature = "vectors")]
    vector_defaults: None,
  }
}

fn request(query: Query, limit: usize, execution: ExecutionStrategy, block: Option<usize>, filter: Option<Filter>) -> SearchRequest {
  SearchRequest {
    query,
    fields: None,
    filter,
    limit,
    return_hits: true,
    candidate_size: None,
    sort: Vec::new(),
    cursor: None,
    execution,
    bmw_block_size: block,
    fuzzy: None,
    #[cfg(feature = "vectors")]
    vector_query: None,
    #[cfg(feature = "vectors")]
    vector_filter: None,
    return_stored: false,
    highlight_field: None,
    highlight: None,
    collapse: None,
    aggs: BTreeMap::new(),
    suggest: BTreeMap::new(),
    rescore: None,
    explain: false,
    profile: false,
  }
}

fn exec_name(e: &ExecutionStrategy) -> &'static str {
  match e {
    ExecutionStrategy::Bm25 => "bm25",
    ExecutionStrategy::Wand => "wand",
    ExecutionStrategy::Bmw => "bmw",
  }
}

fn hits_json(res: &searchlite_core::api::SearchResult) -> Value {
  let hits: Vec<Value> = res
    .hits
    .iter()
    .map(|h| json!({ "doc_id": h.doc_id, "score": h.score, "score_bits": h.score.to_bits(), "vector_score": h.vector_score }))
    .collect();
  json!({ "hits": hits, "total_hits_estimate": res.total_hits_estimate })
}

fn copy_dir(src: &Path, dst: &Path) {
  fs::create_dir_all(dst).unwrap();
  for e in fs::read_dir(src).unwrap() {
    let e = e.unwrap();
    let to = dst.join(e.file_name());
    if e.file_type().unwrap().is_dir() {
      copy_dir(&e.path(), &to);
    } else {
      fs::copy(e.path(), to).unwrap();
    }
  }
}

fn filter_json(f: &Option<Filter>) -> Value {
  match f {
    None => Value::Null,
    Some(f) => serde_json::to_value(f).unwrap(),
  }
}

/// every (query, limit, execution) case of one corpus
fn run_cases(
  reader: &searchlite_core::api::IndexReader,
  cases: &[(Value, Query, usize, Option<Filter>)],
  block: Option<usize>,
) -> Vec<Value> {
  let mut out = Vec::new();
  for (desc, query, limit, filter) in cases {
    for execution in [ExecutionStrategy::Bm25, ExecutionStrategy::Wand, ExecutionStrategy::Bmw] {
      let req = request(query.clone(), *limit, execution.clone(), block, filter.clone());
      let res = reader.search(&req).unwrap();
      let mut case = json!({ "query": desc, "limit": limit, "execution": exec_name(&execution), "bmw_block_size": block,
                             "filter": filter_json(filter) });
      let h = hits_json(&res);
      case["hits"] = h["hits"].clone();
      case["total_hits_estimate"] = h["total_hits_estimate"].clone();
      out.push(case);
    }
  }
  out
}

fn term(field: &str, value: &str) -> QueryNode {
  QueryNode::Term { field: field.into(), value: value.into(), boost: None }
}

fn bool_must(terms: &[&str]) -> QueryNode {
  QueryNode::Bool {
    must: terms.iter().map(|t| term("body", t)).collect(),
    should: vec![],
    must_not: vec![],
    filter: vec![],
    minimum_should_match: None,
    boost: None,
  }
}

#[test]
fn dump_pruning_corpus() {
  // tests/pruning.rs:45-104, same seed, same draw order
  let vocab = ["rust", "search", "engine", "fast", "tiny", "wand", "bmw"];
  let dir = tempfile::tempdir().unwrap();
  let path = dir.path().join("idx");
  let idx = Index::create(&path, Schema::default_text_body(), opts(&path, 1.2, 0.75)).unwrap();
  let mut rng = StdRng::seed_from_u64(42);
  let mut docs = Vec::new();
  {
    let mut writer = idx.writer().unwrap();
    for i in 0..40 {
      let body: Vec<&str> = (0..6).map(|_| vocab[rng.gen_range(0..vocab.len())]).collect();
      let body = body.join(" ");
      writer
        .add_document(&Document {
          fields: [("_id".into(), json!(format!("doc-{i}"))), ("body".into(), json!(body.clone()))].into_iter().collect(),
        })
        .unwrap();
      docs.push(json!({ "id": format!("doc-{i}"), "body": body }));
    }
    writer.commit().unwrap();
  }
  let reader = idx.reader().unwrap();
  let mut cases = Vec::new();
  for _ in 0..5 {
    let mut terms = vocab.to_vec();
    terms.shuffle(&mut rng);
    let q: Vec<&str> = terms.iter().take(3).copied().collect();
    cases.push((json!({ "kind": "query_string", "terms": q }), Query::from(q.join(" ")), 5usize, None));
  }
  let out = json!({ "k1": 1.2, "b": 0.75, "docs": docs, "cases": run_cases(&reader, &cases, Some(4)) });
  fs::write(out_dir().join("ref_pruning.json"), serde_json::to_string_pretty(&out).unwrap()).unwrap();
  copy_dir(&path, &out_dir().join("ref_index_pruning"));
}

#[test]
fn dump_query_ast_corpus() {
  // tests/query_ast.rs:24-66
  let dir = tempfile::tempdir().unwrap();
  let path = dir.path().join("idx");
  let mut schema = Schema::default_text_body();
  schema.keyword_fields.push(KeywordField { name: "lang".into(), stored: true, indexed: true, fast: true, nullable: false });
  schema.numeric_fields.push(NumericField { name: "year".into(), i64: true, fast: true, stored: true, nullable: false });
  let idx = Index::create(&path, schema, opts(&path, 0.9, 0.4)).unwrap();
  let rows = [
    ("doc-1", "rust engine fast", "en", 2024i64),
    ("doc-2", "rust database tiny", "en", 2022),
    ("doc-3", "rust search", "fr", 2021),
    ("doc-4", "rust boring engine", "en", 2020),
    ("doc-5", "fast tiny search", "fr", 2023),
  ];
  let mut docs = Vec::new();
  {
    let mut writer = idx.writer().unwrap();
    for (id, body, lang, year) in rows {
      writer
        .add_document(&Document {
          fields: [("_id".to_string(), json!(id)), ("body".to_string(), json!(body)), ("lang".to_string(), json!(lang)), ("year".to_string(), json!(year))]
            .into_iter()
            .collect::<BTreeMap<_, _>>(),
        })
        .unwrap();
      docs.push(json!({ "id": id, "body": body, "lang": lang, "year": year }));
    }
    writer.commit().unwrap();
  }
  let reader = idx.reader().unwrap();
  let en_recent = Filter::And(vec![
    Filter::KeywordEq { field: "lang".into(), value: "en".into() },
    Filter::I64Range { field: "year".into(), min: 2021, max: 2024 },
  ]);
  let cases = vec![
    (json!({ "kind": "query_string", "terms": ["rust"] }), Query::from("rust"), 10usize, None),
    (json!({ "kind": "query_string", "terms": ["rust", "engine"] }), Query::from("rust engine"), 10, None),
    (json!({ "kind": "query_string", "terms": ["fast", "tiny", "search"] }), Query::from("fast tiny search"), 3, None),
    (json!({ "kind": "bool_must", "terms": ["rust", "engine"] }), Query::from(bool_must(&["rust", "engine"])), 10, None),
    (json!({ "kind": "query_string", "terms": ["rust", "fast"] }), Query::from("rust fast"), 10, Some(en_recent.clone())),
    (json!({ "kind": "bool_must", "terms": ["rust"] }), Query::from(bool_must(&["rust"])), 10, Some(Filter::Not(Box::new(en_recent)))),
  ];
  let out = json!({ "k1": 0.9, "b": 0.4, "docs": docs, "cases": run_cases(&reader, &cases, None) });
  fs::write(out_dir().join("ref_query_ast.json"), serde_json::to_string_pretty(&out).unwrap()).unwrap();
  copy_dir(&path, &out_dir().join("ref_index_query_ast"));
}

/// explicit 64-bit LCG (Knuth MMIX constants): the Python side regenerates nothing — the docs are in the JSON — but the
/// corpus is reproducible from this file alone
struct Lcg(u64);
impl Lcg {
  fn next(&mut self) -> u64 {
    self.0 = self.0.wrapping_mul(6364136223846793005).wrapping_add(1442695040888963407);
    self.0 >> 33
  }
  /// Zipf-ish rank in 0..n: floor(n^u) - 1 with u uniform in (0, 1]
  fn zipf(&mut self, n: usize) -> usize {
    let u = ((self.next() % 1_000_000) as f64 + 1.0) / 1_000_000.0;
    (((n as f64).powf(u)) as usize).clamp(1, n) - 1
  }
}

fn zipf_docs(n_docs: usize, vocab: usize, seed: u64, with_fields: bool) -> Vec<(String, String, String, i64)> {
  let mut rng = Lcg(seed);
  let langs = ["en", "es", "de", "fr"];
  (0..n_docs)
    .map(|i| {
      let len = 8 + (rng.next() % 40) as usize;
      let body: Vec<String> = (0..len).map(|_| format!("w{}", rng.zipf(vocab))).collect();
      let lang = langs[rng.zipf(langs.len())].to_string();
      let year = 2000 + (rng.next() % 26) as i64;
      let _ = with_fields;
      (format!("d{i:06}"), body.join(" "), lang, year)
    })
    .collect()
}

#[test]
fn dump_zipf_corpus() {
  let dir = tempfile::tempdir().unwrap();
  let path = dir.path().join("idx");
  let mut schema = Schema::default_text_body();
  schema.keyword_fields.push(KeywordField { name: "lang".into(), stored: true, indexed: true, fast: true, nullable: false });
  schema.numeric_fields.push(NumericField { name: "year".into(), i64: true, fast: true, stored: true, nullable: false });
  let idx = Index::create(&path, schema, opts(&path, 0.9, 0.4)).unwrap();
  let rows = zipf_docs(600, 120, 20260101, true);
  let mut docs = Vec::new();
  {
    // two commits => two segments: per-segment N / df / avgdl and the SortKey merge are exercised (api/reader.rs:2670-2777)
    for half in rows.chunks(300) {
      let mut writer = idx.writer().unwrap();
      for (id, body, lang, year) in half {
        writer
          .add_document(&Document {
            fields: [("_id".to_string(), json!(id)), ("body".to_string(), json!(body)), ("lang".to_string(), json!(lang)), ("year".to_string(), json!(year))]
              .into_iter()
              .collect::<BTreeMap<_, _>>(),
          })
          .unwrap();
        docs.push(json!({ "id": id, "body": body, "lang": lang, "year": year }));
      }
      writer.commit().unwrap();
    }
  }
  let reader = idx.reader().unwrap();
  let mut rng = Lcg(20260102);
  let mut cases = Vec::new();
  for i in 0..24 {
    let n = 2 + (rng.next() % 4) as usize;
    let mut terms: Vec<String> = Vec::new();
    while terms.len() < n {
      let t = format!("w{}", 2 + rng.zipf(118));
      if !terms.contains(&t) {
        terms.push(t);
      }
    }
    let refs: Vec<&str> = terms.iter().map(|s| s.as_str()).collect();
    let limit = if i % 2 == 0 { 5 } else { 40 };
    cases.push((json!({ "kind": "query_string", "terms": terms }), Query::from(terms.join(" ")), limit, None));
    if i % 3 == 0 {
      cases.push((json!({ "kind": "bool_must", "terms": refs[..2] }), Query::from(bool_must(&refs[..2])), limit, None));
    }
    if i % 4 == 0 {
      let f = Filter::And(vec![
        Filter::KeywordEq { field: "lang".into(), value: "en".into() },
        Filter::I64Range { field: "year".into(), min: 2005, max: 2015 },
      ]);
      cases.push((json!({ "kind": "query_string", "terms": terms }), Query::from(terms.join(" ")), limit, Some(f)));
    }
  }
  let out = json!({ "k1": 0.9, "b": 0.4, "segments": [300, 300], "docs": docs, "cases": run_cases(&reader, &cases, None) });
  fs::write(out_dir().join("ref_zipf.json"), serde_json::to_string_pretty(&out).unwrap()).unwrap();
  copy_dir(&path, &out_dir().join("ref_index_zipf"));
}

#[cfg(feature = "vectors")]
#[test]
fn dump_hybrid_corpus() {
  use searchlite_core::api::types::{LegacyVectorQuery, VectorQuerySpec};
  let dir = tempfile::tempdir().unwrap();
  let path = dir.path().join("idx");
  let schema: Schema = serde_json::from_value(json!({
    "doc_id_field": "_id",
    "text_fields": [ { "name": "body", "analyzer": "default", "stored": true, "indexed": true, "nullable": false } ],
    "keyword_fields": [], "numeric_fields": [], "nested_fields": [],
    "vector_fields": [ { "name": "embedding", "dim": 8, "metric": "Cosine" } ]
  }))
  .unwrap();
  let idx = Index::create(&path, schema, opts(&path, 0.9, 0.4)).unwrap();
  let rows = zipf_docs(200, 60, 20260103, false);
  let mut rng = Lcg(20260104);
  let mut docs = Vec::new();
  {
    let mut writer = idx.writer().unwrap();
    for (i, (id, body, _, _)) in rows.iter().enumerate() {
      let mut fields: BTreeMap<String, Value> = [("_id".to_string(), json!(id)), ("body".to_string(), json!(body))].into_iter().collect();
      let mut emb = Value::Null;
      if i % 9 != 0 {
        // one doc in nine has no vector (missing_vector_score, api/reader.rs:218-223)
        let v: Vec<f32> = (0..8).map(|_| ((rng.next() % 2001) as f32 - 1000.0) / 1000.0).collect();
        emb = json!(v);
        fields.insert("embedding".into(), emb.clone());
      }
      writer.add_document(&Document { fields }).unwrap();
      docs.push(json!({ "id": id, "body": body, "embedding": emb }));
    }
    writer.commit().unwrap();
  }
  let reader = idx.reader().unwrap();
  let mut cases = Vec::new();
  for i in 0..12 {
    let terms: Vec<String> = (0..3).map(|j| format!("w{}", 1 + (i * 3 + j) % 40)).collect();
    let qv: Vec<f32> = (0..8).map(|_| ((rng.next() % 2001) as f32 - 1000.0) / 1000.0).collect();
    for alpha in [0.5f32, 0.2, 0.0] {
      let mut req = request(Query::from(terms.join(" ")), 10, ExecutionStrategy::Bm25, None, None);
      req.candidate_size = Some(200); // every matching doc is a BM25 candidate
      req.vector_query = Some(VectorQuerySpec::Legacy(LegacyVectorQuery("embedding".into(), qv.clone(), alpha)));
      let res = reader.search(&req).unwrap();
      let mut case = json!({ "query": { "kind": "query_string", "terms": terms }, "limit": 10, "execution": "bm25", "candidate_size": 200,
                             "vector": { "field": "embedding", "metric": "cosine", "alpha": alpha, "query_vector": qv } });
      let h = hits_json(&res);
      case["hits"] = h["hits"].clone();
      case["total_hits_estimate"] = h["total_hits_estimate"].clone();
      cases.push(case);
    }
  }
  let out = json!({ "k1": 0.9, "b": 0.4, "docs": docs, "cases": cases,
                    "note": "the reference adds HNSW candidates that match the text query to the BM25 candidates (collect_vector_maps, \
                             api/reader.rs:2379-2475); with candidate_size >= #docs the union equals the BM25 candidate set" });
  fs::write(out_dir().join("ref_hybrid.json"), serde_json::to_string_pretty(&out).unwrap()).unwrap();
  copy_dir(&path, &out_dir().join("ref_index_hybrid"));
}
