--
-- Testing for batching queries
-- Queries with full overlap (subsumption)
--
-- Query in SQL:
-- 10 Queries of the type:
-- SELECT col1 FROM tbl3_batch WHERE col4 >= _ AND col4 < _;
--
--
batch_queries()
s0=select(db1.tbl3_batch.col4,1430,1490)
s1=select(db1.tbl3_batch.col4,1432,1488)
s2=select(db1.tbl3_batch.col4,1434,1486)
s3=select(db1.tbl3_batch.col4,1436,1484)
s4=select(db1.tbl3_batch.col4,1438,1482)
s5=select(db1.tbl3_batch.col4,1440,1480)
s6=select(db1.tbl3_batch.col4,1442,1478)
s7=select(db1.tbl3_batch.col4,1444,1476)
s8=select(db1.tbl3_batch.col4,1446,1474)
s9=select(db1.tbl3_batch.col4,1448,1472)
batch_execute()
f0=fetch(db1.tbl3_batch.col1,s0)
f1=fetch(db1.tbl3_batch.col1,s1)
f2=fetch(db1.tbl3_batch.col1,s2)
f3=fetch(db1.tbl3_batch.col1,s3)
f4=fetch(db1.tbl3_batch.col1,s4)
f5=fetch(db1.tbl3_batch.col1,s5)
f6=fetch(db1.tbl3_batch.col1,s6)
f7=fetch(db1.tbl3_batch.col1,s7)
f8=fetch(db1.tbl3_batch.col1,s8)
f9=fetch(db1.tbl3_batch.col1,s9)
print(f0)
print(f1)
print(f2)
print(f3)
print(f4)
print(f5)
print(f6)
print(f7)
print(f8)
print(f9)
